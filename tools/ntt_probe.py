#!/usr/bin/env python3
"""Where does an NTT pass spend its time?  Runs the trace LDE (2^20 rows x 88 columns -> 2^21, the four digit passes of
ntt_fast.cu + the API's reorder) with the measurement-only instantiations of the 2^10 tile kernels:
  ZKIR_NTT_PROBE=0  the real kernels
  ZKIR_NTT_PROBE=1  memory only   (same loads, twiddle-table reads and stores; no arithmetic, no shared-memory exchange)
  ZKIR_NTT_PROBE=2  arithmetic only (no global loads or table reads; stores predicated off)
If real ~ max(memory, arithmetic) the two already overlap and staging the loads differently (TMA / cp.async.bulk) has nothing
to hide; if real ~ memory + arithmetic they are serialised and a staged pipeline could win up to the smaller of the two.
usage: python tools/ntt_probe.py            (spawns itself once per mode; add `ncu` to get per-launch times)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(cols, log_n, reps):
    import numpy as np
    import zkir_b200
    ctx = zkir_b200.Context(0)
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2013265921, size=(cols, 1 << log_n), dtype=np.uint64).astype(np.uint32)
    d_in = ctx.to_device(a)
    d_out = ctx.alloc(a.nbytes * 2)
    for _ in range(3):
        ctx.lde(d_in, d_out, cols, log_n, 1)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        ctx.lde(d_in, d_out, cols, log_n, 1)
    ms = ctx.timer_stop() / reps
    print(f"probe={os.environ.get('ZKIR_NTT_PROBE', '0')} lde({cols} x 2^{log_n}, blowup 2) incl. reorder: {ms:.4f} ms")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    cols, log_n = 88, 20
    use_ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
    for mode in ("0", "1", "2"):
        env = dict(os.environ, ZKIR_NTT_PROBE=mode)
        cmd = [sys.executable, os.path.abspath(__file__), "child", str(cols), str(log_n), "1" if use_ncu else "10"]
        if use_ncu:
            cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "-k", "regex:dft_tile|coset_reorder", "-s", "15",
                   "--log-file", os.path.join(ROOT, "gpurun_out", f"ntt_probe_{mode}.csv")] + cmd
        subprocess.run(cmd, env=env, check=False)
