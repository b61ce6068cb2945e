#!/usr/bin/env python3
"""Profiler-only numbers for bench.py (DRAM traffic of the LDE launches, pipe utilisation of the hash kernels), extracted from the raw
metric csv of an `ncu --set full` capture (tools/gpu_profile_round.sh) and stamped with the build id of the library they were measured
on.  bench.py quotes them only when that build is the one running.
usage: python tools/ncu_metrics.py <tag> [<build_id>]  ->  profiles/<tag>_ncu_metrics.json"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from build_id import build_id  # noqa: E402

U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "s": 1e3, "second": 1e3}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
                                    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    out = []
    for r in rows[2:]:
        f = lambda n: float(r[ix[n]].replace(",", "")) if r[ix[n]] else 0.0
        out.append({"kernel": r[ix["Kernel Name"]].split("(")[0].replace("void ", ""),
                    "ms": f("gpu__time_duration.sum") * T[units[ix["gpu__time_duration.sum"]]],
                    "dram_read": f("dram__bytes_read.sum") * U[units[ix["dram__bytes_read.sum"]]],
                    "dram_write": f("dram__bytes_write.sum") * U[units[ix["dram__bytes_write.sum"]]],
                    "grid": int(f("launch__grid_size")), "fmaheavy_pct": f("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                    "smem_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), "dram_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
    return out


def main():
    tag = sys.argv[1]
    bid = sys.argv[2] if len(sys.argv) > 2 else build_id()
    G = os.path.join(ROOT, "gpurun_out")
    hot = load(os.path.join(G, f"ncu_hot_{tag}_raw.csv"))
    dft = [k for k in hot if k["kernel"].startswith("dft_tile_kernel<10")]
    big = max(k["grid"] for k in dft)
    # the trace LDE = the four 2^10-digit launches over all 88 columns: the two inverse passes (grid = big / 2) and the two forward passes (2 cosets)
    lde = [k for k in dft if k["grid"] in (big, big // 2)][:4]
    out = {"build_id": bid, "source": f"gpurun_out/ncu_hot_{tag}_raw.csv (ncu --set full --clock-control none, second proof of tools/prove_once.py 2)",
           "lde_launches": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in x.items()} for x in lde],
           "lde_dram_bytes_per_proof": int(sum(k["dram_read"] + k["dram_write"] for k in lde)),
           "lde_ms_under_ncu": sum(k["ms"] for k in lde)}
    for name in ("leaf_hash_rows_kernel", "quotient_kernel", "aux_rows_kernel", "deep_kernel"):
        ks = [k for k in hot if k["kernel"].startswith(name)]
        extra = os.path.join(G, f"ncu_leaf_{tag}_raw.csv")
        if not ks and os.path.exists(extra):
            ks = [k for k in load(extra) if k["kernel"].startswith(name)]
        if ks:
            k = max(ks, key=lambda x: x["ms"])
            out[name.replace("_kernel", "").replace("_rows", "") + "_fmaheavy_active_frac"] = round(k["fmaheavy_pct"] / 100.0, 4)
            out[name.replace("_kernel", "").replace("_rows", "") + "_dram_frac"] = round(k["dram_pct"] / 100.0, 4)
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_metrics.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, json.dumps({k: v for k, v in out.items() if k != "lde_launches"}))


if __name__ == "__main__":
    main()
