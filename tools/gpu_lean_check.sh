#!/bin/bash
# one GPU call: timing of the default and the lean tile kernel (4 / 5 / 6 CTAs per SM), bench-scale parity of the lean kernel,
# the small parity suites under it, ncu --set full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== timing";
timeout 120 python tools/tile_variants.py
for n in 4 5 6; do B200VIS_TILE_KERNEL=lean B200VIS_LEAN_CTAS=$n timeout 120 python tools/tile_variants.py; done
echo "== bench-scale parity (lean)";
timeout 500 python -m pytest tests/test_gpu_bench_scale.py -q -x -m gpu -k "lean" 2>&1 | tail -15
echo "== small suites under lean 5 / 6";
B200VIS_TILE_KERNEL=lean timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -m gpu 2>&1 | tail -4
B200VIS_TILE_KERNEL=lean B200VIS_LEAN_CTAS=6 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -m gpu 2>&1 | tail -4
echo "== ncu lean";
for n in 5 6; do
B200VIS_TILE_KERNEL=lean B200VIS_LEAN_CTAS=$n timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_propagate_cull_lean --launch-skip 40 --launch-count 1 -f -o gpurun_out/r02d_lean$n python tools/tile_variants.py 2>&1 | tail -2
done
} > gpurun_out/lean_check.log 2>&1
tail -40 gpurun_out/lean_check.log
