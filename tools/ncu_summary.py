#!/usr/bin/env python3
"""Markdown summary of an .ncu-rep (ncu --set full): one row per profiled launch with the metrics DESIGN.md quotes."""
import csv, subprocess, sys

WANT = [
    ("time_us", "gpu__time_duration.sum", 1e-3),          # ns -> us (ncu raw csv reports ns or ms depending on version; unit column is read)
    ("dram_rd_MB", "dram__bytes_read.sum", None),
    ("dram_wr_MB", "dram__bytes_write.sum", None),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("fmaheavy_%", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", None),
    ("alu_%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", None),
    ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
    ("regs", "launch__registers_per_thread", None),
    ("grid", "launch__grid_size", None),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", None),
    ("L2_hit_%", "lts__t_sector_hit_rate.pct", None),
]
UNIT_TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
UNIT_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    cols = [(n, hdr.index(m) if m in hdr else -1) for n, m, _ in WANT]
    print("| kernel | " + " | ".join(n for n, _ in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        vals = []
        for n, i in cols:
            if i < 0 or not r[i]:
                vals.append("-"); continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if n == "time_us":
                v *= UNIT_TO_US.get(u, 1.0)
            elif n.endswith("_MB"):
                v *= UNIT_TO_MB.get(u, 1.0)
            vals.append(f"{v:.1f}" if abs(v) < 1e6 else f"{v:.3g}")
        name = r[kn].split("(")[0].replace("void ", "")
        print(f"| `{name[:48]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
