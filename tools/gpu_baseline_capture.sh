#!/bin/bash
# GPU-box script: parity tests, microbenchmarks and ncu --set full captures of the current kernels.
# usage (from the repo root on the box): bash tools/gpu_baseline_capture.sh <tag>
TAG=${1:-cap}
mkdir -p gpurun_out
set -x
./tools/bin/microbench > gpurun_out/microbench_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$TAG.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 30 -c 4 -o gpurun_out/${TAG}_ntt $B > gpurun_out/ncu_ntt_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaf_hash_kernel -c 2 -o gpurun_out/${TAG}_leaf $B > gpurun_out/ncu_leaf_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:quotient_kernel|open_partial_kernel|deep_kernel|leaf_hash_pairs_kernel" -c 5 -o gpurun_out/${TAG}_misc $B > gpurun_out/ncu_misc_$TAG.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/smi_$TAG.txt
nproc >> gpurun_out/smi_$TAG.txt
cat gpurun_out/microbench_$TAG.txt gpurun_out/pytest_$TAG.txt
