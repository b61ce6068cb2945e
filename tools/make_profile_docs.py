#!/usr/bin/env python3
"""Turn the files tools/gpu_profile_round.sh left in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profile_docs.py <tag, e.g. r02v22> <label, e.g. v22> "<one-line description of the build>" [bench json] [ref json]"""
import json
import os
import shutil
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "tools"))
from build_id import build_id  # noqa: E402

tag, label, desc = sys.argv[1], sys.argv[2], sys.argv[3]
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
rnd = tag.split("v")[0]   # r02


def run(*a):
    return subprocess.run([sys.executable] + list(a), capture_output=True, text=True, cwd=R).stdout


bench_path = sys.argv[4] if len(sys.argv) > 4 else None
stage_line = ""
if bench_path:
    bench = json.load(open(bench_path))
    shutil.copy(bench_path, f"{P}/{rnd}_bench_{label}.json")
    if len(sys.argv) > 5:
        shutil.copy(sys.argv[5], f"{P}/{rnd}_bench_ref_{label}.json")
    st = bench["stage_ms"]
    stage_line = (f"\nbench.py stage_ms of the same build (CUDA events on the library stream, {bench['steps']} steps): "
                  + ", ".join(f"{k} {v:.2f}" for k, v in st.items() if k != "h2d") + f" ms; {bench['ms_per_step']:.2f} ms/proof "
                  f"(profiles/{rnd}_bench_{label}.json, build id {bench['config'].get('build_id')}).\n")
head = f"build id {build_id()}"
ls = run("tools/launch_summary.py", f"{G}/launches_{tag}.csv")
open(f"{P}/{rnd}_launches_{label}.md", "w").write(
    f"# {rnd} {label} -- launch list of one 2^20-row proof: {desc} ({head})\n\n"
    "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prove_once.py 2` (second proof; per-launch times are "
    "cold-cache and serialised: read SHARES).\n\n" + ls + stage_line)
bl = run("tools/launch_summary.py", f"{G}/bench_launches_{tag}.csv")
open(f"{P}/{rnd}_bench_launches_{label}.md", "w").write(
    f"# {rnd} {label} -- launch list of the bench command itself: {desc} ({head})\n\n"
    "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --only-headline` "
    "(the last proof inside the 500-launch window; the kernel SHARES must agree with the prove_once list above, not the absolute times).\n\n" + bl)
hot = open(f"{G}/ncu_hot_{tag}.md").read()
tree = open(f"{G}/ncu_tree_{tag}.md").read()
leaf = open(f"{G}/ncu_leaf_{tag}.md").read() if os.path.exists(f"{G}/ncu_leaf_{tag}.md") else ""
open(f"{P}/{rnd}_ncu_{label}.md", "w").write(
    f"# {rnd} {label} -- ncu --set full summaries: {desc} ({head}; one launch per row; second proof of tools/prove_once.py 2)\n\n"
    f"Captured with `bash tools/gpu_profile_round.sh {tag}` (`ncu --set full --clock-control none --import-source on`); the .ncu-rep files are summarised on the "
    "GPU box (this table + the raw metric csv) because gpurun brings back at most 64 MiB.  time_us under ncu is cold-cache and serialised.\n\n"
    + hot + "\n" + tree + ("\n" + leaf if leaf else ""))
if os.path.exists(f"{G}/launches_full_{tag}.csv"):
    fl = run("tools/launch_summary.py", f"{G}/launches_full_{tag}.csv")
    fn = open(f"{G}/ncu_full_{tag}.md").read() if os.path.exists(f"{G}/ncu_full_{tag}.md") else ""
    fo = open(f"{G}/full_once_{tag}.log").read() if os.path.exists(f"{G}/full_once_{tag}.log") else ""
    open(f"{P}/{rnd}_full_profile_{label}.md", "w").write(
        f"# {rnd} {label} -- full AIR profile (248 + 168 columns, all 50 opcodes), mix workload at 2^18 rows ({head})\n\n"
        "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/full_profile_bench.py 18 1` -- the LAST proof of the run, "
        "Program -> Proof (interpreter write log + memory log, device register rebuild + converter `trace_expand_wl_full_kernel`, then the ordinary proof).\n\n" + fl
        + "\nOutput of the same command without the profiler's per-launch serialisation is in the bench JSON (`full_profile_mix`); under ncu:\n\n```\n" + fo + "```\n\n"
        "## ncu --set full of the profile-specific kernels\n\n" + fn)
print("wrote", f"{P}/{rnd}_launches_{label}.md", f"{P}/{rnd}_bench_launches_{label}.md", f"{P}/{rnd}_ncu_{label}.md")
