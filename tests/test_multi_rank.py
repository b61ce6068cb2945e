"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: proof ownership, max-over-ranks timing, digest gather, and the
hand-off of rank 0's communicator id that precedes a sharded proof (the collectives of a sharded proof themselves run inside
the library over NCCL and are covered by tests/test_gpu_sharded.py)."""
import hashlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zkir_b200 import multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_items = 7
        mine = multi.shard_indices(n_items, rank, world)
        local = {i: (b"proof-%d" % i) * (i + 1) for i in mine}          # stand-ins for proof bytes
        mx = multi.max_over_ranks([10.0 + rank, 5.0 - rank])
        sm = multi.sum_over_ranks([float(len(mine))])
        digests = multi.gather_proof_digests(local, n_items)
        # communicator id hand-off: only rank 0 draws an id, every rank ends up with rank 0's 128 bytes
        r, w, ident = multi.exchange_comm_id(None, make_id=lambda: bytes((7 * i + rank) % 256 for i in range(multi.COMM_ID_LEN)))
        assert (r, w) == (rank, world)
        q.put((rank, mine, mx, sm, digests, ident))
    finally:
        dist.destroy_process_group()


def test_world_size_2_bookkeeping():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, mx0, sm0, dg0, id0), (r1, mine1, mx1, sm1, dg1, id1) = out
    assert id0 == id1 == bytes((7 * i) % 256 for i in range(128))        # rank 0's id reached rank 1
    assert mine0 == [0, 2, 4, 6] and mine1 == [1, 3, 5]                   # i = rank (mod world), complete and disjoint
    assert mx0 == mx1 == [11.0, 5.0]                                      # slowest rank bounds the step
    assert sm0 == sm1 == [7.0]
    want = [hashlib.sha256((b"proof-%d" % i) * (i + 1)).digest() for i in range(7)]
    assert dg0 == dg1 == want                                            # every rank sees every proof's digest


def test_single_process_is_identity():
    assert multi.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert multi.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    assert multi.exchange_comm_id(None) == (0, 1, None)                   # no peers: no id is drawn, comm_init is a no-op


def test_shard_plan_partitions_columns_rows_and_planes():
    """The partition of ONE proof over `world` GPUs (zkir_b200_shard_plan, host arithmetic of csrc/prover.cu): over all ranks the
    column ranges, the row segments and the quotient-plane ranges are disjoint and complete, a rank's Merkle leaf segment is its
    row segment times the blowup, and sizes too small to shard fall back to the unsharded plan on every rank alike."""
    import ctypes as C
    import numpy as np
    import zkir_b200
    from zkir_b200 import _ffi
    lib = _ffi.lib()

    def plan(world, rank, cfg, log_n, min_seg=0):
        out = (C.c_uint64 * 8)()
        p = cfg.params()
        assert lib.zkir_b200_shard_plan(world, rank, C.byref(p), log_n, min_seg, out) == 0
        return list(out)

    W = zkir_b200.air_layout.WIDTH
    for world in (2, 4, 8, 64):
        for log_n, log_b in ((20, 1), (24, 1), (16, 2), (12, 1)):
            cfg = zkir_b200.ProverConfig(log_blowup=log_b)
            N, M = 1 << log_n, 1 << (log_n + log_b)
            plans = [plan(world, r, cfg, log_n) for r in range(world)]
            on = plans[0][0]
            assert all(p[0] == on for p in plans)                       # every rank takes the same branch
            if M // world < 4096:
                assert not on                                           # below the segment threshold: whole proof everywhere
            if not on:
                assert all(p[1:] == [0, W, 0, N, 0, 4, M] for p in plans)
                continue
            cols = np.zeros(W, dtype=int); rows = np.zeros(N, dtype=int); planes = np.zeros(4, dtype=int)
            for r, (_, c0, c1, j0, nj, p0, p1, seg) in enumerate(plans):
                cols[c0:c1] += 1; rows[j0:j0 + nj] += 1; planes[p0:p1] += 1
                assert nj * world == N and j0 == r * nj and seg == nj << log_b
                assert c1 - c0 <= -(-W // world) and p1 - p0 <= -(-4 // world)
            assert (cols == 1).all() and (rows == 1).all() and (planes == 1).all()
    # a lower threshold shards small proofs too (what the GPU tests use), and bad arguments are rejected
    cfg = zkir_b200.ProverConfig()
    assert plan(8, 3, cfg, 8, min_seg=2)[0] == 1 and plan(8, 3, cfg, 8)[0] == 0
    out = (C.c_uint64 * 8)()
    p = cfg.params()
    assert lib.zkir_b200_shard_plan(3, 0, C.byref(p), 20, 0, out) == _ffi.ERR_ARG      # world must be a power of two
    assert lib.zkir_b200_shard_plan(4, 4, C.byref(p), 20, 0, out) == _ffi.ERR_ARG      # rank < world
