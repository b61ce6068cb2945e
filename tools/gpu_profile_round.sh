#!/bin/bash
# GPU-box script: launch list of one proof and of the bench command, and ncu --set full of the hot kernels of ONE proof (the second
# one, caches and tables warm).  The .ncu-rep files are summarised ON THE BOX (markdown table + the raw metric csv) and then deleted:
# gpurun only brings back 64 MiB.
# usage: bash tools/gpu_profile_round.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/prove_once.py 2 > gpurun_out/prove_once_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:leaf_hash_rows_kernel|dft_tile_kernel|quotient_kernel|open_partial_kernel|deep_kernel|aux_rows_kernel|scan_write_kernel|io_sum_kernel" -s 26 -c 30 -o gpurun_out/${TAG}_hot python tools/prove_once.py 2 > gpurun_out/ncu_hot_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:compress_kernel|leaf_hash_pairs_kernel|merkle_coop_kernel|fri_fold_kernel" -s 80 -c 14 -o gpurun_out/${TAG}_tree python tools/prove_once.py 2 > gpurun_out/ncu_tree_$TAG.log 2>&1
for r in hot tree; do
  python tools/ncu_summary.py gpurun_out/${TAG}_$r.ncu-rep > gpurun_out/ncu_${r}_$TAG.md
  ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/ncu_${r}_${TAG}_raw.csv
  rm -f gpurun_out/${TAG}_$r.ncu-rep
done
tail -2 gpurun_out/prove_once_$TAG.log
# the full AIR profile (2^18-row mix workload): launch list of the last two proofs (resident columns, then rows through the device
# converter) and ncu --set full of its profile-specific kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full_$TAG.csv python tools/full_profile_bench.py 18 1 > gpurun_out/full_once_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:quotient_kernel_full|aux_rows_kernel_full|trace_expand_full_kernel|trace_expand_wl_full_kernel" -c 8 -o gpurun_out/${TAG}_full python tools/full_profile_bench.py 18 1 > gpurun_out/ncu_full_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/ncu_full_$TAG.md
rm -f gpurun_out/${TAG}_full.ncu-rep
tail -2 gpurun_out/full_once_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/bench_launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --only-headline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
du -sh gpurun_out
