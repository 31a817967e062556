#!/bin/bash
# one GPU call at the end of a change: the whole -m gpu suite, bench.py (N = 1), the ncu launch list of the same command and an
# ncu --set full capture of the tile kernel.  Artefacts land in gpurun_out/ (summaries are copied to profiles/ by hand).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02c}
{
echo "== pytest -m gpu"; date
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -8
echo "== bench"; date
timeout 600 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 600 gpurun_out/${TAG}_bench_n1.json
echo "== reference arm (short)"; date
timeout 300 python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/${TAG}_bench_reference_arm.json
echo "== ncu launch list"; date
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-next-rows > /dev/null 2>&1
echo "== ncu full"; date
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_propagate_cull_lean --launch-skip 40 --launch-count 1 -f -o gpurun_out/${TAG}_lean python tools/tile_variants.py 2>&1 | tail -2
date
} > gpurun_out/${TAG}_check.log 2>&1
tail -50 gpurun_out/${TAG}_check.log
