#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python benchmarks/compare_updates.py > gpurun_out/upd.log 2>&1
cat gpurun_out/upd.log
