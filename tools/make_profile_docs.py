#!/usr/bin/env python3
"""Turn the files tools/gpu_round_*.sh left in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profile_docs.py <tag, e.g. r01v16> <label, e.g. v16> "<one-line description of the build>" """
import csv, json, os, shutil, subprocess, sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, label, desc = sys.argv[1], sys.argv[2], sys.argv[3]
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")


def run(*a):
    return subprocess.run([sys.executable] + list(a), capture_output=True, text=True, cwd=R).stdout


bench = json.load(open(f"{G}/bench_{tag}.json"))
shutil.copy(f"{G}/bench_{tag}.json", f"{P}/r01_bench_{label}.json")
if os.path.exists(f"{G}/bench_ref_{tag}.json"):
    shutil.copy(f"{G}/bench_ref_{tag}.json", f"{P}/r01_bench_ref_{label}.json")
st = bench["stage_ms"]
stage_line = ", ".join(f"{k} {st[k]:.2f}" for k in ("lde", "trace_commit", "quotient", "quotient_commit", "openings", "fri"))
ls = run("tools/launch_summary.py", f"{G}/launches_{tag}.csv")
open(f"{P}/r01_launches_{label}.md", "w").write(
    f"# r01 {label} -- launch list of one 2^20-row proof: {desc} (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n"
    "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prove_once.py 2` (second proof; per-launch times are "
    "cold-cache and serialised: read SHARES).\n\n" + ls +
    f"\nbench.py stage_ms of the same build (CUDA events on the library stream, {bench['steps']} steps): {stage_line} ms; "
    f"{bench['ms_per_step']:.2f} ms/proof (profiles/r01_bench_{label}.json).\n")
hot, tree = run("tools/ncu_summary.py", f"{G}/{tag}_hot.ncu-rep"), run("tools/ncu_summary.py", f"{G}/{tag}_tree.ncu-rep")
open(f"{P}/r01_ncu_{label}.md", "w").write(
    f"# r01 {label} -- ncu --set full summaries: {desc} (one launch per row; second proof of tools/prove_once.py 2)\n\n"
    f"Captured with `bash tools/gpu_profile_round.sh {tag}` (`ncu --set full --clock-control none --import-source on`). Columns as in r01_ncu_v8.md.\n\n"
    + hot + "\n" + tree)
# DRAM traffic of the four trace-LDE launches (first four rows of the hot capture)
out = subprocess.run(["ncu", "-i", f"{G}/{tag}_hot.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
kn, ir, iw, it = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ns": 1e-6, "us": 1e-3, "ms": 1, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6}
L = [{"kernel": r[kn].split("(")[0].replace("void ", "")[:32], "dram_read": float(r[ir].replace(",", "")) * U[units[ir]],
      "dram_write": float(r[iw].replace(",", "")) * U[units[iw]], "time_ms": float(r[it].replace(",", "")) * T.get(units[it], 1e-6)} for r in rows[2:]]
first = next(i for i in range(len(L) - 3) if all("dft_tile_kernel<10" in L[i + j]["kernel"] for j in range(4)))   # the trace LDE: 4 passes in a row
L = L[first:first + 4]
json.dump({"source": f"profiles/r01_ncu_{label}.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the four trace-LDE launches)",
           "width": bench["config"]["width"], "launches": L, "lde_dram_bytes_total_estimate": sum(x["dram_read"] + x["dram_write"] for x in L)},
          open(f"{P}/r01_lde_traffic.json", "w"), indent=1)
# launch list of the bench command
rows = list(csv.reader(open(f"{G}/bench_launches_{tag}.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
recs = [(r[kn].split("(")[0].replace("void ", ""), float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]
tot = sum(c for _, c in recs); agg = {}
for k, c in recs:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += c
o = [f"# r01 {label} -- ncu launch list of the bench command itself: {desc}", "",
     "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
     f"(tools/gpu_bench_launches.sh). {len(recs)} launches, {agg.get('query_kernel', [0])[0]} proofs (device-resident, write-log and full-row end-to-end arms, "
     "two-in-flight arms) + the NTT microbenchmark;", "per-launch times are cold-cache and serialised, so read the SHARES. A number printed by this run is not a bench value.", "",
     "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    o.append(f"| `{k[:60]}` | {v[0]} | {v[1] / 1000:.0f} | {100 * v[1] / tot:.1f}% |")
hs = sum(v[1] for k, v in agg.items() if any(t in k for t in ("leaf_hash", "compress_kernel", "merkle_coop"))) / tot
ns = sum(v[1] for k, v in agg.items() if "dft_tile" in k) / tot
o += ["", f"Hashing kernels {100 * hs:.1f} % of the GPU time of the bench command, NTT tiles {100 * ns:.1f} % (incl. the NTT microbenchmark); "
      f"event-timed stage split of profiles/r01_bench_{label}.json: hashing stages (trace_commit + most of quotient_commit and fri) "
      f"{100 * (st['trace_commit'] + 0.8 * st['quotient_commit'] + 0.85 * st['fri']) / bench['ms_per_step']:.0f} %, lde {100 * st['lde'] / bench['ms_per_step']:.0f} %."]
open(f"{P}/r01_bench_launches_{label}.md", "w").write("\n".join(o) + "\n")
print(f"{label}: {bench['value'] / 1e6:.2f} M cycles/s, {bench['ms_per_step']:.2f} ms/proof, e2e {bench['e2e']['ms_per_step']:.2f} ms, stages: {stage_line}")
print("pipelined", bench["pipelined"]["ms_per_proof"], bench["pipelined"].get("e2e_ms_per_proof"), "cpu", bench.get("cpu_baseline", {}).get("value"))
print("roofline", bench["roofline"]["achieved"], bench["roofline"]["frac"], "ntt", bench["ntt_roofline"]["achieved"], bench["ntt_roofline"]["frac"])
