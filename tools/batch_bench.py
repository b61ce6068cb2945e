#!/usr/bin/env python3
"""BASELINE config 4: many independent tiny proofs (a+b, 11 cycles -> 2^4 rows).  Times zkir_b200_prove_batch."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkir_b200

SRC = """
addi r10, r0, 1
ecall
add r1, r10, r0
addi r10, r0, 1
ecall
add r11, r1, r10
addi r10, r0, 2
ecall
addi r10, r0, 0
addi r11, r0, 0
ecall
"""
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prog = zkir_b200.assemble(SRC)
cfg = zkir_b200.ProverConfig()
traces = []
for i in range(n):
    res = zkir_b200.VM(prog, [i, 2 * i + 1], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    assert res.outputs == [3 * i + 1]
    traces.append(res.pack())
ctx = zkir_b200.Context(0)
ctx.prove_batch([c for c, _ in traces[:8]], [p for _, p in traces[:8]], cfg)
t0 = time.perf_counter()
proofs = ctx.prove_batch([c for c, _ in traces], [p for _, p in traces], cfg)
dt = time.perf_counter() - t0
ok = all(zkir_b200.verify(pb, cfg, pv)[0] for pb, (_, pv) in zip(proofs[:16], traces[:16]))
print(f"{n} tiny proofs: {dt * 1e3:.1f} ms total, {dt / n * 1e3:.3f} ms/proof, {n / dt:.0f} proofs/s, verified first 16: {ok}")
