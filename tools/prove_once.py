#!/usr/bin/env python3
"""Profiling driver: build the 2^20-row fibonacci trace once, prove it `reps` times from device memory."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkir_b200
from zkir_b200.workloads import fib_trace, fib_program_input

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 209715
use_writelog = len(sys.argv) > 3 and sys.argv[3] == "writelog"   # no host packing: needed for very long traces
ctx = zkir_b200.Context(0)
cfg = zkir_b200.ProverConfig()
if use_writelog:
    t0 = time.perf_counter()
    res = zkir_b200.VM(fib_program_input(), [n], zkir_b200.VMConfig(max_cycles=1 << 26, enable_execution_trace=True)).run()
    wl = res.writelog()
    log_n = res.min_log_n()
    print(f"VM + write log: {time.perf_counter() - t0:.2f} s, {res.cycles} cycles, log_n={log_n}")
    for i in range(reps):
        t0 = time.perf_counter()
        pb, pv = ctx.prove_writelog(wl, cfg, log_n)
        dt = (time.perf_counter() - t0) * 1e3
        print(f"proof {i}: {dt:.2f} ms wall, stages {ctx.stage_ms()}")
else:
    res, cols, pv = fib_trace(n_input=n)
    log_n = int(cols.shape[1]).bit_length() - 1
    ctx.set_program(res)
    d = ctx.to_device(cols)
    for i in range(reps):
        t0 = time.perf_counter()
        pb = ctx.prove_columns(None, pv, cfg, device_resident=(d, log_n))
        dt = (time.perf_counter() - t0) * 1e3
        print(f"proof {i}: {dt:.2f} ms wall, stages {ctx.stage_ms()}")
ok, why = zkir_b200.verify(pb, cfg, pv, res)
print("verify", ok, why)
