#!/usr/bin/env python3
"""ONE proof sharded over the GPUs of a box (BASELINE config 5), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_prove.py [log_n ...]            (default: 20 24)

Every rank builds the same fibonacci write log, proves it once unsharded (the reference bytes), links the contexts with
`Context.comm_init` (NCCL inside the library) and proves it `reps` times collectively.  Rank 0 prints one JSON line per size:
single-GPU and sharded ms/proof (CUDA events on the library's stream, max over ranks), stage times of rank 0, and whether the
sharded proof bytes equal the single-GPU bytes on every rank and pass the CPU verifier."""
import json
import os
import sys
import numpy as np  # noqa: F401

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import zkir_b200  # noqa: E402
from zkir_b200 import multi  # noqa: E402
from conftest import fib_program_input  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [20, 24]
    reps = int(os.environ.get("REPS", "5"))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zkir_b200.Context(local)
    cfg = zkir_b200.ProverConfig()
    for log_n in sizes:
        n = ((1 << log_n) + 2) // 5                      # 5n - 2 cycles <= 2^log_n
        res = zkir_b200.VM(fib_program_input(), [n], zkir_b200.VMConfig(max_cycles=1 << 26, enable_execution_trace=True)).run()
        wl0 = res.writelog()
        assert res.min_log_n() == log_n, (res.cycles, log_n)
        pins = {k: zkir_b200.PinnedBuffer(wl0[k].shape, wl0[k].dtype) for k in ("pcs", "instrs", "wlog")}   # the hand-off buffers are pinned
        wl = dict(wl0)
        for k, pb in pins.items():
            pb.array[...] = wl0[k]
            wl[k] = pb.array
        ctx.comm_shutdown()
        single = []
        for _ in range(3):
            ctx.timer_start()
            want, pv = ctx.prove_writelog(wl, cfg, log_n)
            single.append(ctx.timer_stop())
        single_stage = ctx.stage_ms()
        if world > 1:
            dist.barrier()
        ctx.comm_init()
        ms, same = [], True
        for _ in range(reps + 2):
            if world > 1:
                dist.barrier()
            ctx.timer_start()
            got, pv2 = ctx.prove_writelog(wl, cfg, log_n)
            ms.append(ctx.timer_stop())
            same = same and got == want
        stage = ctx.stage_ms()
        t = multi.max_over_ranks([min(single), sorted(ms[2:])[len(ms[2:]) // 2], 0.0 if same else 1.0], device="cuda" if world > 1 else "cpu")
        if rank == 0:
            ok, why = zkir_b200.verify(got, cfg, pv2)
            print(json.dumps({"config": f"fibonacci 2^{log_n}-row trace, one proof sharded over {world} GPU(s)", "cycles": res.cycles, "n_gpus": world,
                              "single_gpu_ms": round(t[0], 3), "sharded_ms": round(t[1], 3), "speedup": round(t[0] / t[1], 3),
                              "cycles_per_s_sharded": round(res.cycles / t[1] * 1e3), "bytes_identical_on_all_ranks": t[2] == 0.0,
                              "verifier_accepts": bool(ok), "stage_ms_single": {k: round(v, 3) for k, v in single_stage.items()},
                              "stage_ms_sharded_rank0": {k: round(v, 3) for k, v in stage.items()}}), flush=True)
        del res, wl, wl0
        for pb in pins.values():
            pb.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
