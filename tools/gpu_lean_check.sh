#!/bin/bash
# one GPU call: lean tile kernel with the top levels walked in registers (default) vs through the level loop (B200VIS_LEAN_PROBE=4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== timing";
export B200VIS_TILE_KERNEL=lean B200VIS_LEAN_CTAS=4
B200VIS_LEAN_PROBE=4 timeout 120 python tools/tile_variants.py
timeout 120 python tools/tile_variants.py
B200VIS_LEAN_PROBE=4 timeout 120 python tools/tile_variants.py
timeout 120 python tools/tile_variants.py
echo "== parity (lean4)";
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -m gpu 2>&1 | tail -4
unset B200VIS_TILE_KERNEL B200VIS_LEAN_CTAS
timeout 400 python -m pytest tests/test_gpu_bench_scale.py -q -x -m gpu -k "lean" 2>&1 | tail -5
echo "== ncu lean4";
B200VIS_TILE_KERNEL=lean B200VIS_LEAN_CTAS=4 timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_propagate_cull_lean --launch-skip 40 --launch-count 1 -f -o gpurun_out/r02f_lean4top python tools/tile_variants.py 2>&1 | tail -2
} > gpurun_out/lean_check.log 2>&1
tail -40 gpurun_out/lean_check.log
