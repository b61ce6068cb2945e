#!/bin/bash
# GPU-box script: ncu --set full of the hot kernels of ONE proof (the second one, caches and tables warm) + launch list.
# usage: bash tools/gpu_profile_round.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/prove_once.py 2 > gpurun_out/prove_once_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:leaf_hash_kernel|dft_tile_kernel|quotient_kernel|open_partial_kernel|deep_kernel|trace_expand" -s 16 -c 18 -o gpurun_out/${TAG}_hot python tools/prove_once.py 2 > gpurun_out/ncu_hot_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:compress_kernel|leaf_hash_pairs_kernel|merkle_coop_kernel" -s 58 -c 12 -o gpurun_out/${TAG}_tree python tools/prove_once.py 2 > gpurun_out/ncu_tree_$TAG.log 2>&1
tail -2 gpurun_out/prove_once_$TAG.log
