#!/bin/bash
# round-1 v13 evidence: full bench (N=1), launch lists, ncu --set full of the hot kernels
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01v17.json 2> gpurun_out/bench_r01v17.err
tail -c 400 gpurun_out/bench_r01v17.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01v17.json 2>/dev/null
bash tools/gpu_profile_round.sh r01v17 > /dev/null 2>&1
bash tools/gpu_bench_launches.sh r01v17 > /dev/null 2>&1
python tools/ntt_bench.py > gpurun_out/ntt_bench_r01v17.txt 2>&1
ls -la gpurun_out | tail -15
