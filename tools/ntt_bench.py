#!/usr/bin/env python3
"""Device timing of the NTT / LDE entry points (CUDA events on torch's current stream are useless here: the library
launches on its own stream, so time with wall clock around ctx.sync())."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkir_b200

ctx = zkir_b200.Context(0)
rng = np.random.default_rng(1)
for log_n, cols in [(20, 32), (20, 85), (16, 85)]:
    a = rng.integers(0, 2013265921, size=(cols, 1 << log_n), dtype=np.uint64).astype(np.uint32)
    d = ctx.to_device(a)
    for _ in range(3):
        ctx.ntt(d, cols, log_n)
    ctx.sync()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        ctx.ntt(d, cols, log_n)
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    print(f"ntt natural 2^{log_n} x {cols}: {ms:.3f} ms  {8 * (1 << log_n) * cols / ms / 1e6:.0f} GB/s (8nC)")
    d_out = ctx.alloc(cols * (8 << log_n))
    for _ in range(3):
        ctx.lde(d, d_out, cols, log_n, 1)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.lde(d, d_out, cols, log_n, 1)
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    print(f"lde(+reorder) 2^{log_n} x {cols}: {ms:.3f} ms  {12 * (1 << log_n) * cols / ms / 1e6:.0f} GB/s (4NC(1+B))")
    ctx.free(d); ctx.free(d_out)
