#!/usr/bin/env python3
"""How much does a second in-flight proof (own context + stream + host thread on the SAME GPU) fill the latency-bound phases?"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import zkir_b200
from conftest import fib_trace

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
res, cols, pv = fib_trace(n_input=209715)
log_n = 20
cfg = zkir_b200.ProverConfig()
for nctx in (1, 2, 3):
    ctxs = [zkir_b200.Context(0) for _ in range(nctx)]
    devs = [c.to_device(cols) for c in ctxs]
    for c, d in zip(ctxs, devs):
        c.prove_columns(None, pv, cfg, device_resident=(d, log_n))
        c.prove_columns(None, pv, cfg, device_resident=(d, log_n))
    def work(c, d):
        for _ in range(K):
            c.prove_columns(None, pv, cfg, device_resident=(d, log_n))
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(c, d)) for c, d in zip(ctxs, devs)]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{nctx} context(s): {nctx * K} proofs in {dt * 1e3:.1f} ms -> {dt / (nctx * K) * 1e3:.2f} ms/proof, {nctx * K * res.cycles / dt / 1e6:.1f} M cycles/s")
    for c, d in zip(ctxs, devs):
        c.free(d); c.close()
