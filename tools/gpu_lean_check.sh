#!/bin/bash
# one GPU call: the lean tile kernel with drifting warps (PIPE) -- smoke, parity, timing against B200VIS_LEAN_PIPE=0, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
export B200VIS_TILE_KERNEL=lean
echo "== smoke";
if ! timeout 90 python -c "import __graft_entry__ as g; g.smoke()"; then echo "SMOKE FAILED rc=$?"; exit 0; fi
echo "== small suites";
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -m gpu --timeout 60 2>&1 | tail -6
echo "== timing";
B200VIS_LEAN_PIPE=0 timeout 120 python tools/tile_variants.py
timeout 120 python tools/tile_variants.py
B200VIS_LEAN_PIPE=0 timeout 120 python tools/tile_variants.py
timeout 120 python tools/tile_variants.py
unset B200VIS_TILE_KERNEL
echo "== bench-scale parity";
timeout 400 python -m pytest tests/test_gpu_bench_scale.py -q -x -m gpu -k "lean" 2>&1 | tail -5
echo "== ncu";
B200VIS_TILE_KERNEL=lean timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_propagate_cull_lean --launch-skip 40 --launch-count 1 -f -o gpurun_out/r02g_pipe python tools/tile_variants.py 2>&1 | tail -2
} > gpurun_out/lean_check.log 2>&1
tail -40 gpurun_out/lean_check.log
