#!/usr/bin/env python3
"""Stage times of one full-profile proof (mix workload, 2^18 rows by default), trace resident.  ZKIR_QUOTIENT_VARIANT selects the quotient kernel shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkir_b200
from zkir_b200.workloads import mix_program

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5      # profiling runs pass 1: one warm proof, one measured, one from rows
iters = ((1 << log_n) - 16) // 22
res = zkir_b200.VM(mix_program(), [iters], zkir_b200.VMConfig(max_cycles=1 << 26, enable_execution_trace=True)).run()
cols, pv = res.pack()
ctx = zkir_b200.Context(0)
ctx.set_program(res)
d = ctx.to_device(cols)
cfg = zkir_b200.ProverConfig()
ln = int(cols.shape[1]).bit_length() - 1
for _ in range(2 if reps > 1 else 1):
    pb = ctx.prove_columns(cols, pv, cfg, device_resident=(d, ln))
assert zkir_b200.verify(pb, cfg, pv, res) == (True, "")
ctx.timer_start()
for _ in range(reps):
    ctx.prove_columns(cols, pv, cfg, device_resident=(d, ln))
ms = ctx.timer_stop() / reps
print(f"variant={os.environ.get('ZKIR_QUOTIENT_VARIANT', 'default')} rows=2^{ln} cycles={res.cycles} ms/proof={ms:.3f} " + " ".join(f"{k}={v:.3f}" for k, v in ctx.stage_ms().items()))
pb_rows, _ = ctx.prove_rows(res.rows(), cfg, ln, profile="full")   # the device converter + the same proof
assert pb_rows == pb
print("from rows (host memory replay + device converter): stage_ms", {k: round(v, 3) for k, v in ctx.stage_ms().items()})
import time
proof = zkir_b200.prove(res.program, [iters], cfg)    # Program -> Proof: write log + memory log, no host replay
assert proof.bytes_ == pb
t0 = time.perf_counter()
for _ in range(reps):
    proof = zkir_b200.prove(res.program, [iters], cfg)
print(f"Program -> Proof (interpreter + write log + memory log + device converter): {(time.perf_counter() - t0) / reps * 1e3:.3f} ms wall, stage_ms",
      {k: round(v, 3) for k, v in proof.stage_ms.items()})
