#!/usr/bin/env python3
"""bench.py -- headline benchmark of the trace->proof hot path (BASELINE.json metric).

Step = one full proof of the 2^20-row fibonacci trace (BASELINE configs[1]): LDE, Merkle commitments, quotient,
openings, FRI, queries.  `value` = real VM cycles proved per second with the packed trace already in HBM
(device time of the step from CUDA events on the library's launch stream); `e2e` = the same through
zkir_b200_prove_writelog with the interpreter's register write log (pc, word, written register and value per cycle) in
pinned HOST memory: the H2D copy, the device-side register reconstruction + converter and the D2H of the proof are inside
the timed region (wall clock bracketed by device syncs); `e2e.full_rows` is the same with the reference's full TraceRow data.  N>1: every rank proves its own trace (independent proofs, no
data-path collective; weak scaling), value = cycles of all ranks / max-over-ranks time.

`--impl reference`: the reference contains no prover, so the reference arm times this
repo's CPU oracle (oracle/, kind "port") on all host cores on the SAME workload (same trace, AIR and parameters); the number
of timed proofs is cut to a time budget and the line says how many were run.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "cycles_proved_per_sec"
UNIT = "cycles/s"
FIB_N_FULL = 209715          # 5n-2 = 1_048_573 cycles -> 2^20 rows
CPU_BUDGET_S = 200.0         # the CPU arm proves the SAME workload; the number of timed proofs is cut to fit this budget


def workload_name(fib_n, log_n, width, cfg):
    return (f"fibonacci via input tape n={fib_n} ({5 * fib_n - 2} cycles -> 2^{log_n}-row trace, W={width}, "
            f"log_blowup={cfg.log_blowup}, {cfg.num_queries} queries, {cfg.pow_bits} PoW bits)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """One long-lived `nvidia-smi -lms 100` child (the recipe's clocks line); rows are read as they stream in."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
        except Exception:
            pass

    def mark(self):
        return len(self.rows)

    def summary(self, first=0):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        rows = self.rows[first:] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(rows), "window": "device-resident + end-to-end timed loops"}


def make_trace(fib_n):
    """-> (ExecutionResult, cols, pv, interpreter seconds, host packer seconds)"""
    from zkir_b200.workloads import timed_fib_trace
    return timed_fib_trace(fib_n)


def cpu_oracle_run(fib_n, steps, warmup, budget_s=CPU_BUDGET_S):
    """Time the CPU oracle prover (all host cores, OpenMP) on the SAME workload as the GPU arm.  Only the checker lives in
    oracle/; this is one of the two places allowed to execute it (cpu_baseline / --impl reference).  `steps` is cut so that the
    run fits `budget_s` (the first proof is the probe); returns what was really run."""
    from oracle.binding import Oracle
    import zkir_b200
    res, cols, pv, _, _ = make_trace(fib_n)
    o = Oracle()
    # all host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1); `cores` = the team size in effect
    cores = int(o.l.oracle_set_threads(len(os.sched_getaffinity(0))))
    cfg = zkir_b200.ProverConfig()
    t0 = time.time()
    pb = o.prove(cfg, cols, pv, res)           # probe (and the one warm-up proof when warmup >= 1)
    probe = time.time() - t0
    ok, why = zkir_b200.verify(pb, cfg, pv, res)
    if not ok:
        raise SystemExit(f"CPU oracle proof rejected by the verifier: {why}")
    warm_done = 1
    for _ in range(max(0, min(warmup, 1) - warm_done)):
        o.prove(cfg, cols, pv, res)
    if warmup >= 1:
        steps = max(1, min(steps, int(budget_s / max(probe, 1e-3))))
        t0 = time.time()
        for _ in range(steps):
            o.prove(cfg, cols, pv, res)
        dt = (time.time() - t0) / steps
    else:                                  # no warm-up asked: the probe IS the single timed proof
        steps, dt = 1, probe
    log_n = int(cols.shape[1]).bit_length() - 1
    return {"value": res.cycles / dt, "s_per_proof": dt, "cores": cores, "cycles": res.cycles, "log_n": log_n, "steps": steps,
            "warmup": min(warmup, 1), "width": int(cols.shape[0]), "cfg": cfg}


def ncu_metrics(build):
    """Numbers that only a profiler can give (DRAM traffic, pipe utilisation) live in profiles/*_ncu_metrics.json, written by
    tools/ncu_metrics.py from a capture of a named build; they are quoted only when that build is the one running."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)):
        if f.endswith("_ncu_metrics.json"):
            try:
                j = json.load(open(os.path.join(pdir, f)))
            except Exception:
                continue
            if j.get("build_id") == build:
                best = dict(j, file="profiles/" + f)
    return best


_JSON_FD = None


def emit_json(obj):
    """The ONE line of the contract goes to the process's original stdout; everything else any library prints (NCCL's version banner
    and NCCL_DEBUG output go to fd 1) was redirected to stderr by main()."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fib-n", type=int, default=FIB_N_FULL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the sections for BASELINE configs 3, 4 and 5")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_oracle_run(args.fib_n, args.steps, args.warmup)
        v = r["value"]
        sample = (f"the full workload: fibonacci n={args.fib_n}, {r['cycles']} cycles, 2^{r['log_n']} rows, same AIR and parameters as the GPU arm; "
                  f"{r['steps']} timed proofs of {r['s_per_proof']:.2f} s after {r['warmup']} warm-up (asked: --steps {args.steps} --warmup {args.warmup}, cut to a {CPU_BUDGET_S:.0f} s budget)")
        emit_json({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"],
            "ms_per_step": r["s_per_proof"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (BabyBear mod p)",
            "data": "synthetic",
            "config": {"workload": workload_name(args.fib_n, r["log_n"], r["width"], r["cfg"]), "rows": 1 << r["log_n"], "width": r["width"],
                       "reference_note": "seceq/zkir contains no prover; the CPU arm is this repo's oracle (own restatement of docs/PROVER_SPEC.md, OpenMP, not Plonky3)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return 0

    import torch
    import zkir_b200
    from zkir_b200 import _ffi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the proving path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    from build_id import build_id
    build = build_id()
    res, cols, pv, vm_s, pack_s = make_trace(args.fib_n)
    cycles = res.cycles
    log_n = int(cols.shape[1]).bit_length() - 1
    ctx = zkir_b200.Context(local_rank)
    ctx.set_program(res)
    cfg = zkir_b200.ProverConfig()
    workload = workload_name(args.fib_n, log_n, int(cols.shape[0]), cfg)
    # the interpreter's raw rows (TraceRow: pc, word, 16 registers) in pinned host memory: what crosses PCIe per step
    rows = res.rows()
    pin = {k: zkir_b200.PinnedBuffer(rows[k].shape, rows[k].dtype) for k in ("pcs", "instrs", "regs")}
    for k in pin:
        pin[k].array[...] = rows[k]
    prows = dict(rows, **{k: pin[k].array for k in pin})
    rows_bytes = int(sum(pin[k].array.nbytes for k in pin))
    # the same execution as a register write log (pc32, word, (reg << 56 | value)): 16 B/row, also pinned
    T = int(rows["pcs"].shape[0])
    pin_wl = {"pcs": zkir_b200.PinnedBuffer((T,), np.uint32), "wlog": zkir_b200.PinnedBuffer((T,), np.uint64), "instrs": pin["instrs"]}
    wl = res.writelog(out={"pcs": pin_wl["pcs"].array, "wlog": pin_wl["wlog"].array})
    wl["instrs"] = pin["instrs"].array
    wl_bytes = int(wl["pcs"].nbytes + wl["wlog"].nbytes + wl["instrs"].nbytes)
    d_trace = ctx.to_device(cols)
    proof_bytes = 0

    # ---------------- device-resident arm (value)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        pb = ctx.prove_columns(None, pv, cfg, device_resident=(d_trace, log_n))
    proof_bytes = len(pb)
    ok, why = zkir_b200.verify(pb, cfg, pv, res)
    if not ok:
        raise SystemExit(f"proof rejected by the verifier: {why}")
    barrier()
    clk_mark = sampler.mark()
    l0 = ctx.kernel_launches
    dev_ms, stage_acc = 0.0, {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.prove_columns(None, pv, cfg, device_resident=(d_trace, log_n))
        st = ctx.stage_ms()
        dev_ms += sum(st.values())
        for k, v in st.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.kernel_launches - l0
    # ---------------- end-to-end arm (host buffers through the C ABI)
    for _ in range(2):
        pb_rows, pv_rows = ctx.prove_rows(prows, cfg, log_n)
        pb_wl, pv_wl = ctx.prove_writelog(wl, cfg, log_n)
    if pb_rows != pb or list(pv_rows) != list(pv) or pb_wl != pb or list(pv_wl) != list(pv):
        raise SystemExit("device converters (rows / write log) and the host converter disagree")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.prove_writelog(wl, cfg, log_n)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_stage = ctx.stage_ms()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.prove_rows(prows, cfg, log_n)
    barrier()
    e2e_rows_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.summary(clk_mark)

    # ---------------- throughput mode: two proofs in flight on this GPU (second context = own stream + host thread); the
    # latency-bound phases of one proof (FRI, Merkle tops, Fiat-Shamir steps) overlap the hashing of the other
    ctx2 = zkir_b200.Context(local_rank)
    ctx2.set_program(res)
    d_trace2 = ctx2.to_device(cols)
    for _ in range(2):
        ctx2.prove_columns(None, pv, cfg, device_resident=(d_trace2, log_n))

    def _stream_of_proofs(c, d):
        for _ in range(args.steps):
            c.prove_columns(None, pv, cfg, device_resident=(d, log_n))

    barrier()
    t0 = time.perf_counter()
    workers = [threading.Thread(target=_stream_of_proofs, args=a) for a in ((ctx, d_trace), (ctx2, d_trace2))]
    for t in workers:
        t.start()
    for t in workers:
        t.join()
    barrier()
    pipe_ms = (time.perf_counter() - t0) * 1e3
    # the same two-in-flight mode END TO END: both contexts prove from the pinned write log (H2D, converter and D2H inside)
    ctx2.prove_writelog(wl, cfg, log_n)

    def _stream_of_e2e_proofs(c):
        for _ in range(args.steps):
            c.prove_writelog(wl, cfg, log_n)

    barrier()
    t0 = time.perf_counter()
    workers = [threading.Thread(target=_stream_of_e2e_proofs, args=(c,)) for c in (ctx, ctx2)]
    for t in workers:
        t.start()
    for t in workers:
        t.join()
    barrier()
    pipe_e2e_ms = (time.perf_counter() - t0) * 1e3
    # ... and Program -> Proof in that mode: while one context's GPU work runs, the other thread's interpreter executes the next program
    from zkir_b200.workloads import fib_program_input
    cfg_prog = zkir_b200.ProverConfig(max_cycles=cycles + 16)
    prog_in = fib_program_input()
    for c in (ctx, ctx2):
        c.prove_program(prog_in, [args.fib_n], cfg_prog)

    def _stream_of_programs(c):
        for _ in range(args.steps):
            c.prove_program(prog_in, [args.fib_n], cfg_prog)

    barrier()
    t0 = time.perf_counter()
    workers = [threading.Thread(target=_stream_of_programs, args=(c,)) for c in (ctx, ctx2)]
    for t in workers:
        t.start()
    for t in workers:
        t.join()
    barrier()
    pipe_prog_ms = (time.perf_counter() - t0) * 1e3
    ctx2.free(d_trace2)
    ctx2.close()

    # ---------------- NTT roofline microbench: one forward transform of the trace size on all W columns (8*n*C algorithmic
    # bytes), natural order in and out, timed with CUDA events on the library's stream
    ntt_log, ntt_cols = log_n, int(cols.shape[0])
    rng = np.random.default_rng(0x5EED)
    d_ntt = ctx.to_device(rng.integers(0, 2013265921, size=(ntt_cols, 1 << ntt_log), dtype=np.uint64).astype(np.uint32))
    for _ in range(3):
        ctx.ntt(d_ntt, ntt_cols, ntt_log)
    ctx.sync()
    reps = 10
    ctx.timer_start()
    for _ in range(reps):
        ctx.ntt(d_ntt, ntt_cols, ntt_log)
    ntt_ms = ctx.timer_stop() / reps
    ctx.free(d_ntt)
    ctx.free(d_trace)

    # ---------------- Program -> Proof in ONE call (zkir_b200_prove_program): the interpreter records the register write log into
    # pinned memory while it runs, chunks are uploaded during the run, then rebuild + convert + prove.  Same proof bytes required.
    for _ in range(2):
        pb_prog, pv_prog, cyc_prog, ln_prog = ctx.prove_program(prog_in, [args.fib_n], cfg_prog)
    if pb_prog != pb or cyc_prog != cycles or list(pv_prog) != list(pv):
        raise SystemExit("zkir_b200_prove_program and the step-by-step path disagree")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.prove_program(prog_in, [args.fib_n], cfg_prog)
    barrier()
    prog_ms = (time.perf_counter() - t0) * 1e3
    prog_stage = ctx.stage_ms()

    # ---------------- the other BASELINE configs, a few repetitions each (skipped with --only-headline)
    extra = {}
    if not args.only_headline:
        from zkir_b200.workloads import pos2_program, pos2_cycles, POS2_ITERS_FULL, add_program
        # config 3: SYS_POSEIDON2 loop, 2^18 - 1 permutations, 16 cycles each -> 2^22 rows.  The interpreter executes the permutations
        # (host); the AIR constrains the syscall row as a register write only (no memory argument / hash chip yet: docs/PROVER_SPEC.md 3.5)
        cfg3 = zkir_b200.ProverConfig(max_cycles=pos2_cycles(POS2_ITERS_FULL) + 16, enable_poseidon2_syscall=True)
        p3 = pos2_program()
        vm = zkir_b200.VM(p3, [POS2_ITERS_FULL], zkir_b200.VMConfig(max_cycles=cfg3.max_cycles, enable_poseidon2_syscall=True))
        t0 = time.perf_counter(); r3 = vm.run(); c3_vm_s = time.perf_counter() - t0       # interpreter alone (no recording): also yields the public I/O transcript
        pb3, pv3, cyc3, ln3 = ctx.prove_program(p3, [POS2_ITERS_FULL], cfg3)
        ok3, why3 = zkir_b200.verify(pb3, cfg3, pv3, p3, io=r3.io)
        if not ok3 or cyc3 != pos2_cycles(POS2_ITERS_FULL):
            raise SystemExit(f"config 3 proof rejected: {why3}")
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            ctx.prove_program(p3, [POS2_ITERS_FULL], cfg3)
        barrier()
        c3_ms = (time.perf_counter() - t0) * 1e3 / 2
        c3_stage = ctx.stage_ms()
        extra["config3_poseidon2_loop"] = {
            "workload": f"SYS_POSEIDON2 loop, {POS2_ITERS_FULL} permutations x 16 cycles = {cyc3} cycles -> 2^{ln3}-row trace, per GPU",
            "program_to_proof_ms": c3_ms, "value": world * cyc3 / (c3_ms * 1e-3), "unit": UNIT, "interpreter_only_s": c3_vm_s,
            "proof_stage_ms": {k: v for k, v in c3_stage.items() if k != "h2d"}, "proof_ms": sum(v for k, v in c3_stage.items() if k != "h2d"),
            "note": "Program -> Proof wall time; the interpreter (host, one thread) executes the permutations and dominates; the syscall is constrained as a register write only"}
        # config 4: a + b, 11 cycles -> 2^10 rows, 4096 independent proofs over all GPUs (zkir_b200_prove_batch: 8 worker contexts per GPU)
        n4 = 4096 // world
        pa = add_program()
        ctx.set_program(pa)
        traces = []
        for i in range(n4):
            g = rank * n4 + i
            r4 = zkir_b200.VM(pa, [g, 2 * g + 1], zkir_b200.VMConfig(enable_execution_trace=True)).run()
            traces.append(r4.pack() + (r4.io,))
        cfg4 = zkir_b200.ProverConfig()
        ctx.prove_batch([c for c, _, _ in traces[:64]], [q for _, q, _ in traces[:64]], cfg4, io_list=[e for _, _, e in traces[:64]])     # warm: worker contexts, tables, graphs
        barrier()
        t0 = time.perf_counter()
        out4 = ctx.prove_batch([c for c, _, _ in traces], [q for _, q, _ in traces], cfg4, io_list=[e for _, _, e in traces])
        barrier()
        c4_ms = (time.perf_counter() - t0) * 1e3
        for i in (0, n4 // 2, n4 - 1):
            ok4, why4 = zkir_b200.verify(out4[i], cfg4, traces[i][1], pa, io=traces[i][2])
            if not ok4:
                raise SystemExit(f"config 4 proof {i} rejected: {why4}")
        extra["config4_add_batch"] = {"workload": f"a + b (11 cycles -> 2^10 rows), 4096 independent proofs over {world} GPU(s), {n4} per GPU", "batch_ms": c4_ms,
                                      "proofs_per_s": 4096 / (c4_ms * 1e-3), "proof_bytes": len(out4[0])}
        ctx.set_program(res)
        del traces, out4
        # full AIR profile (docs/PROVER_SPEC.md 3.6-3.8: all 50 opcodes; 248 main + 168 aux columns): a loop of MUL / MULH / DIVU / REMU,
        # shifts, bitwise operations, signed compares and 8-byte stores / loads, 2^18 rows.  The interpreter records full rows, the
        # host packer builds the wide table (outside the timed region); timed = the proof from device-resident columns.
        from zkir_b200.workloads import mix_program, mix_cycles
        it_f = ((1 << 18) - 16) // 22
        pf = mix_program()
        t0 = time.perf_counter()
        rf = zkir_b200.VM(pf, [it_f], zkir_b200.VMConfig(max_cycles=1 << 20, enable_execution_trace=True)).run()
        f_vm_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        cols_f, pv_f = rf.pack()
        f_pack_s = time.perf_counter() - t0
        ln_f = int(cols_f.shape[1]).bit_length() - 1
        ctx.set_program(rf)
        d_f = ctx.to_device(cols_f)
        cfg_f = zkir_b200.ProverConfig()
        pb_f = ctx.prove_columns(cols_f, pv_f, cfg_f, device_resident=(d_f, ln_f))
        ok_f, why_f = zkir_b200.verify(pb_f, cfg_f, pv_f, rf)
        if not ok_f:
            raise SystemExit(f"full-profile proof rejected: {why_f}")
        barrier()
        ctx.timer_start()
        for _ in range(3):
            ctx.prove_columns(cols_f, pv_f, cfg_f, device_resident=(d_f, ln_f))
        f_ms = ctx.timer_stop() / 3
        f_stage = ctx.stage_ms()
        ctx.free(d_f)
        # end to end from the recorded rows in host memory (zkir_b200_prove_rows, full width): host replay of the run's memory, H2D of the rows
        # (152 B/row), device converter, proof, D2H; and Program -> Proof including the interpreter (zkir_b200_prove_program)
        rows_f = rf.rows()
        pb_r, pv_r = ctx.prove_rows(rows_f, cfg_f, ln_f, profile="full")
        if pb_r != pb_f:
            raise SystemExit("full profile: zkir_b200_prove_rows and the packed-columns path disagree")
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.prove_rows(rows_f, cfg_f, ln_f, profile="full")
        f_rows_ms = (time.perf_counter() - t0) * 1e3 / 3
        f_conv_ms = ctx.stage_ms()["h2d"]
        cfg_fp = zkir_b200.ProverConfig(max_cycles=1 << 20)
        pb_p, _, cyc_p, _ = ctx.prove_program(pf, [it_f], cfg_fp)
        if pb_p != pb_f or cyc_p != rf.cycles:
            raise SystemExit("full profile: zkir_b200_prove_program and the packed-columns path disagree")
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.prove_program(pf, [it_f], cfg_fp)
        f_prog_ms = (time.perf_counter() - t0) * 1e3 / 3
        # end to end from the write log + memory log in PINNED host memory (zkir_b200_prove_writelog_mem, 28 B/row): the full profile's
        # counterpart of the headline `e2e`
        cap_f = int(rf.cycles) + 8
        pin_f = {k: zkir_b200.PinnedBuffer((cap_f,), dt) for k, dt in (("wlog", np.uint64), ("mem_old", np.uint64), ("pcs", np.uint32), ("instrs", np.uint32), ("mem_pts", np.uint32))}
        out_f = {k: b.array for k, b in pin_f.items()}
        out_f["mem_old"][:] = 0
        out_f["mem_pts"][:] = 0
        wl_f = zkir_b200.VM(pf, [it_f], zkir_b200.VMConfig(max_cycles=cap_f)).run_writelog(out_f, memory_log=True).writelog()
        pb_w, _ = ctx.prove_writelog(wl_f, cfg_f, ln_f)
        if pb_w != pb_f:
            raise SystemExit("full profile: zkir_b200_prove_writelog_mem and the packed-columns path disagree")
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.prove_writelog(wl_f, cfg_f, ln_f)
        f_wl_ms = (time.perf_counter() - t0) * 1e3 / 3
        f_wl_conv_ms = ctx.stage_ms()["h2d"]
        del wl_f, out_f
        for b in pin_f.values():
            b.close()
        extra["full_profile_mix"] = {
            "workload": f"MUL/MULH/DIVU/REMU + shifts + bitwise + SLT + SD/LD loop, {rf.cycles} cycles -> 2^{ln_f}-row trace, full AIR profile ({cols_f.shape[0]} main columns), per GPU",
            "ms_per_proof": f_ms, "value": world * rf.cycles / (f_ms * 1e-3), "unit": UNIT, "proof_bytes": len(pb_f),
            "stage_ms": {k: v for k, v in f_stage.items() if k != "h2d"}, "interpreter_full_rows_s": f_vm_s, "host_packer_s": f_pack_s,
            "e2e_rows": {"ms_per_proof": f_rows_ms, "value": world * rf.cycles / (f_rows_ms * 1e-3), "h2d_and_convert_ms": f_conv_ms, "h2d_bytes_per_step": int(rf.cycles) * 152,
                         "api": "zkir_b200_prove_rows (full width): TraceRow data in host memory -> host replay of the run's memory -> device converter -> proof bytes"},
            "e2e": {"ms_per_proof": f_wl_ms, "value": world * rf.cycles / (f_wl_ms * 1e-3), "h2d_and_convert_ms": f_wl_conv_ms, "h2d_bytes_per_step": int(rf.cycles) * 28,
                    "d2h_bytes_per_step": len(pb_f),
                    "api": "zkir_b200_prove_writelog_mem: register write log + memory log in pinned host memory -> device register rebuild + 248-column converter -> proof bytes in host memory"},
            "program_to_proof": {"ms_per_proof": f_prog_ms, "value": world * rf.cycles / (f_prog_ms * 1e-3), "api": "zkir_b200_prove_program (full width), interpreter included: register write log + memory log (28 B/cycle) in pinned memory, chunks uploaded while it runs, device converter"},
            "note": "`value`: trace resident in HBM (packed by the host packer outside the timed region, which the e2e paths do not use)"}
        ctx.set_program(res)
        del cols_f

    # ---------------- N > 1 only: ONE proof sharded over the N GPUs (BASELINE config 5 mode; the headline `value` stays N
    # independent proofs).  Collective zkir_b200_prove_writelog from pinned host memory: column-sharded LDE with NVLink row
    # scatter, row-sharded hashing / quotient / DEEP, NCCL for segment roots and query pieces.  Same proof bytes required.
    shard_ms, shard_stage, shard_sha, shard24, shard24_ms = 0.0, {}, None, None, 0.0
    if dist is not None:
        import hashlib
        ctx.comm_init()
        for _ in range(2):
            pb_sh, _ = ctx.prove_writelog(wl, cfg, log_n)
        # ENFORCED on every rank: the sharded proof must be the single-GPU proof, byte for byte, and the verifier must accept it.
        # A mismatch ends the benchmark with a non-zero exit code (no JSON line), it is not a reported flag.
        shard_sha = hashlib.sha256(pb_sh).hexdigest()
        ok_sh, why_sh = zkir_b200.verify(pb_sh, cfg, pv, res)
        same = torch.tensor([1 if (pb_sh == pb and ok_sh) else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        print(f"[rank {rank}] one_proof_sharded sha256={shard_sha} single_gpu sha256={hashlib.sha256(pb).hexdigest()} verifier={'ok' if ok_sh else why_sh}",
              file=sys.stderr, flush=True)
        if int(same.item()) != 1:
            raise SystemExit(f"[rank {rank}] sharded proof differs from the single-GPU proof (or was rejected by the verifier) on at least one rank")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.prove_writelog(wl, cfg, log_n)
        barrier()
        shard_ms = (time.perf_counter() - t0) * 1e3
        shard_stage = ctx.stage_ms()
        if world == 8 and not args.only_headline:
            # config 5: ONE 2^24-row trace sharded over the 8 GPUs (fibonacci n = 3355443: 16_777_213 cycles), Program -> log on the
            # host once (outside the timed region), then collective zkir_b200_prove_writelog from pinned memory
            n24 = 3355443
            vm24 = zkir_b200.VM(fib_program_input(), [n24], zkir_b200.VMConfig(max_cycles=1 << 24))
            T24 = 5 * n24 - 2
            pins = {"pcs": zkir_b200.PinnedBuffer((1 << 24,), np.uint32), "instrs": zkir_b200.PinnedBuffer((1 << 24,), np.uint32), "wlog": zkir_b200.PinnedBuffer((1 << 24,), np.uint64)}
            r24 = vm24.run_writelog({k: v.array for k, v in pins.items()})
            wl24 = r24.writelog()
            pb24, pv24 = ctx.prove_writelog(wl24, cfg, 24)
            ok24, why24 = zkir_b200.verify(pb24, cfg, pv24, r24) if rank == 0 else (True, "")
            sha24 = hashlib.sha256(pb24).hexdigest()
            shas = [None] * world
            dist.all_gather_object(shas, sha24)
            if not ok24 or len(set(shas)) != 1 or r24.cycles != T24:
                raise SystemExit(f"[rank {rank}] config 5: sharded 2^24 proof rejected or ranks disagree: {why24} {shas}")
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                ctx.prove_writelog(wl24, cfg, 24)
            barrier()
            shard24_ms = (time.perf_counter() - t0) * 1e3 / 2
            shard24 = {"workload": f"fibonacci n={n24}: {T24} cycles -> ONE 2^24-row trace row/column-sharded over 8 GPUs", "cycles": T24, "unit": UNIT,
                       "stage_ms_rank0": ctx.stage_ms(), "proof_sha256_all_ranks": sha24, "verifier": "accepted (rank 0)", "proof_bytes": len(pb24)}
            del pins, wl24
        ctx.comm_shutdown()

    # max over ranks
    c3v = extra.get("config3_poseidon2_loop", {}).get("program_to_proof_ms", 0.0)
    c4v = extra.get("config4_add_batch", {}).get("batch_ms", 0.0)
    vals = torch.tensor([dev_ms, wall_ms, e2e_ms, e2e_rows_ms, pipe_ms, shard_ms, pipe_e2e_ms, vm_s, prog_ms, c3v, c4v, shard24_ms, pipe_prog_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, e2e_rows_ms, pipe_ms, shard_ms, pipe_e2e_ms, vm_s, prog_ms, c3v, c4v, shard24_ms, pipe_prog_ms = [float(x) for x in vals.tolist()]
    if "config3_poseidon2_loop" in extra:
        extra["config3_poseidon2_loop"].update(program_to_proof_ms=c3v, value=world * pos2_cycles(POS2_ITERS_FULL) / (c3v * 1e-3))
    if "config4_add_batch" in extra:
        extra["config4_add_batch"].update(batch_ms=c4v, proofs_per_s=4096 / (c4v * 1e-3))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    K = args.steps
    peak, peak_src = peaks()
    N, B, W = 1 << log_n, 1 << cfg.log_blowup, cols.shape[0]
    lde_bytes = 4 * N * W * (1 + B)
    lde_ms = stage_acc["lde"] / K
    lde_gbs = lde_bytes / (lde_ms * 1e-3) / 1e9
    ntt_gbs = 8 * (1 << ntt_log) * ntt_cols / (ntt_ms * 1e-3) / 1e9
    commit_ms = stage_acc["trace_commit"] / K
    leaf_rows = 2 * B                                   # rows per Merkle leaf (docs/PROVER_SPEC.md 4.1)
    hash_perms = B * N * ((W + 7) // 8) + (B * N // leaf_rows - 1)   # leaf sponge absorptions + tree compressions of the trace commitment
    hash_gps = hash_perms / (commit_ms * 1e-3) / 1e9
    prof = ncu_metrics(build)   # profiler-only numbers, quoted only if they were captured on THIS build
    lde_traffic = prof.get("lde_dram_bytes_per_proof") if prof else None
    out = {
        "metric": METRIC, "value": world * cycles / (dev_ms / K * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": dev_ms / K, "wall_ms_per_step": wall_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 (BabyBear mod p, Montgomery)", "data": "synthetic",
        "config": {"workload": workload, "rows": N, "width": int(W), "log_blowup": cfg.log_blowup, "num_queries": cfg.num_queries,
                   "pow_bits": cfg.pow_bits, "l2": f"inputs exceed L2 (trace {4 * N * W >> 20} MiB, LDE {4 * N * W * B >> 20} MiB per step)", "parallelism": f"{world} independent proofs (one per GPU)",
                   "vm_trace_seconds": round(vm_s, 4), "proof_bytes": proof_bytes, "build_id": build},
        "clocks": clocks,
        "e2e": {"value": world * cycles / (e2e_ms / K * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": wl_bytes, "d2h_bytes_per_step": proof_bytes, "h2d_and_convert_ms": e2e_stage["h2d"],
                "api": "zkir_b200_prove_writelog: the interpreter's register write log (pc, word, reg<<56|value) in pinned host memory -> proof bytes in host memory",
                "program_to_proof": {"value": world * cycles / (prog_ms / K * 1e-3), "unit": UNIT, "ms_per_step": prog_ms / K, "h2d_and_convert_ms": prog_stage.get("h2d"),
                                     "api": "zkir_b200_prove_program: Program + inputs -> proof bytes; the interpreter (host, one thread) records the write log into pinned memory, chunks are uploaded while it runs",
                                     "note": "cycles/s INCLUDING the interpreter run (SURVEY.md 8d: 'also report including VM')"},
                "interpreter": {"full_rows_seconds": vm_s, "note": "VM::run with the reference's full TraceRow recording (pc, word, 16 registers per cycle), for comparison"},
                "full_rows": {"value": world * cycles / (e2e_rows_ms / K * 1e-3), "ms_per_step": e2e_rows_ms / K, "h2d_bytes_per_step": rows_bytes,
                              "api": "zkir_b200_prove_rows: TraceRow data as recorded upstream (pc, word, regs[16])"}},
        "pipelined": {"in_flight_per_gpu": 2, "value": world * 2 * K * cycles / (pipe_ms * 1e-3), "unit": UNIT, "ms_per_proof": pipe_ms / (2 * K),
                      "e2e_value": world * 2 * K * cycles / (pipe_e2e_ms * 1e-3), "e2e_ms_per_proof": pipe_e2e_ms / (2 * K),
                      "program_to_proof_value": world * 2 * K * cycles / (pipe_prog_ms * 1e-3), "program_to_proof_ms_per_proof": pipe_prog_ms / (2 * K),
                      "note": "throughput mode: two contexts (stream + host thread each) per GPU, trace resident (`value`) and end to end from the pinned write log (`e2e_value`); the headline `value` / `e2e` above are the one-proof-at-a-time numbers"},
        "gpu_launches": int(launches),
        "stage_ms": {k: v / K for k, v in stage_acc.items()},
        "roofline": {"kernel": f"dft_tile_kernel: trace LDE stage = 2 inverse + 2 forward (x{B} cosets) digit passes (radix-32 register tiles) over {W} columns, 2^{log_n} -> 2^{log_n + cfg.log_blowup} points",
                     "bound": "hbm", "achieved": lde_gbs, "peak": peak, "unit": "GB/s", "frac": lde_gbs / peak, "traffic": lde_traffic,
                     "traffic_source": (prof or {}).get("file"), "algorithmic_bytes": lde_bytes, "ms": lde_ms, "peak_source": peak_src},
        "ntt_roofline": {"kernel": f"dft_tile_kernel x2: forward NTT 2^{ntt_log} x {ntt_cols} columns via zkir_b200_ntt, natural order in/out (8*n*C bytes)", "bound": "hbm",
                         "achieved": ntt_gbs, "peak": peak, "unit": "GB/s", "frac": ntt_gbs / peak, "ms": ntt_ms, "timing": "CUDA events on the library stream, 10 launches"},
        "hash_roofline": {"kernel": f"leaf_hash_rows_kernel (Poseidon2 sponge over the {B * N} LDE rows of the {W}-column trace, {(W + 7) // 8} permutations per row, {leaf_rows} rows per leaf) + Merkle levels",
                          "bound": "integer-multiply pipe", "permutations": hash_perms,
                          "achieved": hash_gps, "unit": "G permutations/s", "ms": commit_ms,
                          "ncu_fmaheavy_active_frac": (prof or {}).get("leaf_hash_fmaheavy_active_frac"), "ncu_source": (prof or {}).get("file") if (prof or {}).get("leaf_hash_fmaheavy_active_frac") is not None else None,
                          "share_of_step": commit_ms / (dev_ms / K)},
    }
    if world > 1:
        out["one_proof_sharded"] = {
            "ms_per_proof": shard_ms / K, "value": cycles / (shard_ms / K * 1e-3), "unit": UNIT, "n_gpus": world,
            "speedup_vs_one_gpu_e2e": (e2e_ms / K) / (shard_ms / K),
            "proof_bytes_identical_to_single_gpu": True, "identity_check": "enforced on every rank before timing (bytes == single-GPU proof, verifier accepts; a mismatch exits non-zero)",
            "proof_sha256": shard_sha, "stage_ms_rank0": shard_stage, "h2d_bytes_per_step_per_gpu": wl_bytes // world,
            "api": "zkir_b200_comm_init + collective zkir_b200_prove_writelog (end to end from pinned host memory, max over ranks)"}
    out.update(extra)
    if shard24 is not None:
        out["config5_sharded_2p24"] = dict(shard24, ms_per_proof=shard24_ms, value=shard24["cycles"] / (shard24_ms * 1e-3))
    if not args.no_cpu_baseline:
        r = cpu_oracle_run(args.fib_n, 1, 0)    # the same workload, proved once (about 10 s on 16 host threads)
        out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                               "sample": f"CPU oracle (OpenMP, {r['cores']} threads) proving the full workload once: fibonacci n={args.fib_n}, {r['cycles']} cycles, 2^{r['log_n']} rows ({r['s_per_proof']:.2f} s)"}
    emit_json(out)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
