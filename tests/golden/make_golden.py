#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json from this repo's oracle (the reference has no vectors for the proving
half: SURVEY.md 8c).  Run from the repo root: `python tests/golden/make_golden.py`."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zkir_b200  # noqa: E402
from conftest import Oracle, fib_trace  # noqa: E402

P = 2013265921
o = Oracle()
g27 = pow(31, 15, P)
gold = {"roots": [pow(g27, 1 << (27 - k), P) for k in range(28)]}
gold["poseidon2_of_0_to_15"] = o.poseidon2(np.arange(16, dtype=np.uint32).reshape(1, 16))[0].tolist()
gold["poseidon2_of_zeros"] = o.poseidon2(np.zeros((1, 16), dtype=np.uint32))[0].tolist()
res, cols, pv = fib_trace(30)
cfg = zkir_b200.ProverConfig(num_queries=20, pow_bits=8)
pb = o.prove(cfg, cols, pv, res)
gold["fib30_trace_sha256"] = hashlib.sha256(cols.tobytes()).hexdigest()
gold["fib30_proof_sha256"] = hashlib.sha256(pb).hexdigest()
gold["fib30_proof_len"] = len(pb)
# full AIR profile (docs/PROVER_SPEC.md 3.6-3.8): the mix workload (multiplier, shifts, bitwise, signed compare, 8-byte stores / loads), 40 iterations
from zkir_b200.workloads import mix_program  # noqa: E402
from zkir_b200.runtime import FULL_WIDTH  # noqa: E402
rf = zkir_b200.VM(mix_program(), [40], zkir_b200.VMConfig(enable_execution_trace=True)).run()
cf, pvf = rf.pack()
pbf = Oracle(width=FULL_WIDTH).prove(cfg, cf, pvf, rf)
gold["mix40_full_trace_sha256"] = hashlib.sha256(cf.tobytes()).hexdigest()
gold["mix40_full_proof_sha256"] = hashlib.sha256(pbf).hexdigest()
gold["mix40_full_proof_len"] = len(pbf)
json.dump(gold, open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w"), indent=1)
print("wrote golden.json", gold["fib30_proof_sha256"])
