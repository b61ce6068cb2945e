#!/bin/bash
# ncu launch list of the bench command itself (shares of the step; cold-cache, serialised times)
TAG=${1:-bench}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
tail -c 600 gpurun_out/bench_under_ncu_$TAG.log
