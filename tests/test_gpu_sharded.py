"""GPU tests of ONE proof sharded over several GPUs (BASELINE config 5; include/zkir_b200.h: zkir_b200_comm_*).

The bar is the same as for the single-GPU path: proof BYTES identical -- to the single-GPU proof (and through it to the
oracle, tests/test_gpu_parity.py).  On a one-GPU box the sharded code path (segment leaf hashing, segment subtrees in the
global tree layout, row/column-sharded sweeps, owner-only query pieces) is exercised by `emulate_shards`: the one context
computes every segment in turn and no collective is needed because the segments share its memory.  With two or more GPUs the
same path runs with one context per GPU and NCCL in between."""
import ctypes as C
import threading

import numpy as np
import pytest

import zkir_b200
from conftest import fib_trace

pytestmark = pytest.mark.gpu


def _first_diff(got, want):
    g, w = np.frombuffer(got, dtype=np.uint32), np.frombuffer(want, dtype=np.uint32)
    return int(np.nonzero(g != w)[0][0])


@pytest.mark.parametrize("n,log_n,log_b,nq,shards,min_seg", [
    (30, None, 1, 20, 2, 2), (30, None, 1, 20, 8, 2), (30, None, 2, 10, 4, 4),
    (205, None, 1, 100, 4, 2), (1000, None, 1, 30, 8, 64), (1000, None, 3, 12, 2, 8), (300, 12, 1, 20, 16, 2),
    (3000, 14, 1, 40, 8, 0), (3000, 16, 1, 100, 8, 0), (3000, 16, 2, 25, 4, 0), (3000, 17, 1, 16, 64, 0),
])
def test_emulated_shards_give_single_gpu_proof_bytes(gpu_ctx, n, log_n, log_b, nq, shards, min_seg):
    res, cols, pv = fib_trace(n, log_n=log_n)
    cfg = zkir_b200.ProverConfig(log_blowup=log_b, num_queries=nq, pow_bits=6)
    want = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    ctx = zkir_b200.Context(0)
    try:
        ctx.emulate_shards(shards, min_seg)
        got = ctx.prove_columns(cols, pv, cfg, program=res)
    finally:
        ctx.close()
    assert len(got) == len(want)
    if got != want:
        pytest.fail(f"sharded proof differs from the single-GPU proof first at word {_first_diff(got, want)}")
    ok, why = zkir_b200.verify(got, cfg, pv, res)
    assert ok, why


def test_emulated_shards_full_size(gpu_ctx):
    """BASELINE config 2 size (2^20 rows), 8 segments, default thresholds: same bytes as the unsharded proof."""
    res, cols, pv = fib_trace(n_input=209715)
    cfg = zkir_b200.ProverConfig()
    gpu_ctx.set_program(res)
    d = gpu_ctx.to_device(cols)
    want = gpu_ctx.prove_columns(None, pv, cfg, device_resident=(d, 20))
    gpu_ctx.emulate_shards(8)
    try:
        got = gpu_ctx.prove_columns(None, pv, cfg, device_resident=(d, 20))
    finally:
        gpu_ctx.emulate_shards(1)
        gpu_ctx.free(d)
    assert got == want


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n,log_n,world", [(30, None, 2), (3000, 16, 2), (3000, 18, 4), (3000, 18, 8)])
def test_nccl_sharded_proof_equals_single_gpu_proof(gpu_ctx, monkeypatch, n, log_n, world):
    """One context per GPU (threads of this process stand in for the processes), NCCL between them: every rank returns the
    single-GPU proof bytes."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    res, cols, pv = fib_trace(n, log_n=log_n)
    cfg = zkir_b200.ProverConfig(num_queries=50, pow_bits=8)
    want = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    lib = zkir_b200._ffi.lib()
    ident = (C.c_uint8 * 128)()
    assert lib.zkir_b200_comm_unique_id(ident) == 0, lib.zkir_b200_last_error(None)
    out, errs = [None] * world, [None] * world

    def rank_main(r):
        try:
            ctx = zkir_b200.Context(r)
            try:
                ctx._check(lib.zkir_b200_comm_init(ctx._h, ident, r, world))
                out[r] = [ctx.prove_columns(cols, pv, cfg, program=res) for _ in range(2)]
                ctx.comm_shutdown()
            finally:
                ctx.close()
        except Exception as e:  # noqa: BLE001
            errs[r] = e

    if n == 30:
        monkeypatch.setenv("ZKIR_SHARD_MIN_SEG", "2")   # read by comm_init: shard even the tiny trees
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank hung"
    assert errs == [None] * world, errs
    for r in range(world):
        assert out[r][0] == want and out[r][1] == want, f"rank {r} differs at word {_first_diff(out[r][0], want)}"


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_sharded_full_profile_proof_equals_single_gpu_proof(gpu_ctx, world):
    """The full AIR profile (248 + 168 columns) over real NCCL: packed columns, recorded rows (host memory replay + device converter on
    every rank) and Program -> Proof (the sharded context takes the rows path) all return the single-GPU proof bytes on every rank."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from zkir_b200.workloads import mix_program
    prog, inputs = mix_program(), [2900]                     # 63 811 cycles -> 2^16 rows
    res = zkir_b200.VM(prog, inputs, zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    rows = res.rows()
    cfg = zkir_b200.ProverConfig(num_queries=50, pow_bits=8, max_cycles=1 << 17)
    want = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    assert zkir_b200.verify(want, cfg, pv, res) == (True, "")
    lib = zkir_b200._ffi.lib()
    ident = (C.c_uint8 * 128)()
    assert lib.zkir_b200_comm_unique_id(ident) == 0, lib.zkir_b200_last_error(None)
    out, errs = [None] * world, [None] * world

    def rank_main(r):
        try:
            ctx = zkir_b200.Context(r)
            try:
                ctx._check(lib.zkir_b200_comm_init(ctx._h, ident, r, world))
                a = ctx.prove_columns(cols, pv, cfg, program=res)
                b = ctx.prove_rows(rows, cfg)[0]
                c = ctx.prove_program(prog, inputs, cfg)[0]
                out[r] = [a, b, c]
                ctx.comm_shutdown()
            finally:
                ctx.close()
        except Exception as e:  # noqa: BLE001
            errs[r] = e

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank hung"
    assert errs == [None] * world, errs
    for r in range(world):
        for k, name in enumerate(("columns", "rows", "program")):
            assert out[r][k] == want, f"rank {r} {name} differs at word {_first_diff(out[r][k], want)}"
