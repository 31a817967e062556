#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 60 python -m pytest tests/test_gpu_bench_scale.py -q -x -m gpu -k "config3_bench and default and not serial" 2>&1 | tail -3
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -m gpu 2>&1 | tail -3
B200VIS_LEAN_PROBE=8 timeout 40 python tools/tile_variants.py
timeout 40 python tools/tile_variants.py
timeout 60 python -m pytest tests/test_gpu_bench_scale.py -q -x -m gpu -k "config4 and default" 2>&1 | tail -3
} > gpurun_out/last.log 2>&1
cat gpurun_out/last.log
