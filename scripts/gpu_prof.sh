#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/prof_bench.log 2>&1
tail -3 gpurun_out/prof_bench.log
wc -l gpurun_out/launches.csv
