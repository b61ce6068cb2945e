#!/bin/bash
# one-proof-over-N-GPUs measurements (run with gpurun --gpus 8): NCCL tests at world 2/4/8, then the sharded prover at N = 2, 4, 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k nccl 2>&1 | tail -5 | tee gpurun_out/sharded_pytest.txt
: > gpurun_out/sharded_r01v14.jsonl
for n in 2 4 8; do
  sizes="20 22"; [ "$n" = 8 ] && sizes="20 22 24"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/sharded_prove.py $sizes 2>&1 \
    | grep "^{\|rror\|Traceback" | tee -a gpurun_out/sharded_r01v14.jsonl
done
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
# the bench contract at N = 8 (headline = 8 independent proofs; one_proof_sharded = the same GPUs on one proof)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | grep "^{" > gpurun_out/bench_8gpu_r01v14.json
tail -c 1500 gpurun_out/bench_8gpu_r01v14.json
