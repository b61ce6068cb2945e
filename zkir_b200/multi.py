"""Multi-GPU plumbing for the trace->proof path: one process per GPU (torch.distributed, NCCL on the GPU box, gloo in CPU tests).

The path shards by INDEPENDENT PROOFS (BASELINE config 4: `i = rank (mod world)`; bench.py N>1: every rank proves its own
trace): there is no data-path collective, only the bookkeeping below -- which proofs a rank owns, the max-over-ranks time the
benchmark contract asks for, and the gather of per-proof digests so rank 0 can report/verify the whole batch.
ONE proof can also be sharded over the GPUs (BASELINE config 5, `Context.comm_init`): the library then runs NCCL collectives
itself (all-gather of Merkle segment roots, all-reduce of the query pieces); the only host-side step is handing rank 0's
128-byte communicator id to the other ranks, `exchange_comm_id` below.  DESIGN.md section 5."""
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


COMM_ID_LEN = 128


def broadcast_bytes(payload, n, src=0, group=None, device="cpu"):
    """Rank `src` passes `payload` (n bytes); every rank returns those n bytes."""
    buf = torch.zeros(n, dtype=torch.uint8, device=device)
    if payload is not None:
        assert len(payload) == n
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(buf, src=src, group=group)
    return bytes(buf.cpu().tolist())


def exchange_comm_id(lib, rank=None, world=None, group=None, make_id=None):
    """Rank 0 draws the communicator id (zkir_b200_comm_unique_id, or `make_id()` in CPU tests), everyone receives it.
    Returns (rank, world, id bytes)."""
    import ctypes as C
    inited = dist.is_available() and dist.is_initialized()
    if rank is None:
        rank = dist.get_rank(group) if inited else 0
    if world is None:
        world = dist.get_world_size(group) if inited else 1
    ident = None
    if world > 1 and rank == 0:
        if make_id is not None:
            ident = make_id()
        else:
            raw = (C.c_uint8 * COMM_ID_LEN)()
            rc = lib.zkir_b200_comm_unique_id(raw)
            if rc != 0:
                raise RuntimeError("zkir_b200_comm_unique_id: " + lib.zkir_b200_last_error(None).decode())
            ident = bytes(raw)
    if world > 1:
        dev = "cuda" if inited and dist.get_backend(group) == "nccl" else "cpu"
        ident = broadcast_bytes(ident, COMM_ID_LEN, 0, group, dev)
    return rank, world, ident
