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
