"""Multi-GPU plumbing for the trace->proof path: one process per GPU (torch.distributed, NCCL on the GPU box, gloo in CPU tests).

The path shards by INDEPENDENT PROOFS (BASELINE config 4: `i = rank (mod world)`; bench.py N>1: every rank proves its own
trace): there is no data-path collective, only the bookkeeping below -- which proofs a rank owns, the max-over-ranks time the
benchmark contract asks for, and the gather of per-proof digests so rank 0 can report/verify the whole batch.
DESIGN.md section 5 explains what row-sharding ONE proof would additionally need."""
import hashlib

import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Proof i is owned by rank i % world."""
    return list(range(rank, n_items, world))


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (device timings -> the slowest rank bounds the job)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def gather_proof_digests(local, n_items, device="cpu"):
    """local: {proof index: proof bytes} of this rank.  Returns on every rank the list of SHA-256 digests (32 bytes each)
    of all n_items proofs, so that any rank can check the batch is complete and identical to a single-GPU run."""
    buf = torch.zeros((n_items, 32), dtype=torch.uint8, device=device)
    for i, pb in local.items():
        buf[i] = torch.tensor(list(hashlib.sha256(pb).digest()), dtype=torch.uint8, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        wide = buf.to(torch.int32)
        dist.all_reduce(wide, op=dist.ReduceOp.SUM)   # rows are disjoint across ranks: SUM == gather
        buf = wide.to(torch.uint8)
    return [bytes(buf[i].tolist()) for i in range(n_items)]
