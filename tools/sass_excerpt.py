#!/usr/bin/env python3
"""SASS evidence for DESIGN.md: per-kernel mnemonic histogram (which pipe the instructions go to) and the head of the listing,
from the library that is actually shipped (cuobjdump -sass on zkir_b200/libzkir_b200.so; no GPU needed).
usage: python tools/sass_excerpt.py <tag>   ->  profiles/<tag>_sass_<kernel>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from build_id import build_id  # noqa: E402

KERNELS = {
    "leaf_hash_rows_kernel": r"leaf_hash_rows_kernel",
    "dft_tile_10_mode0_fwd": r"dft_tile_kernelILi10ELi0ELb0ELi10ELb0ELi0E",
    "dft_tile_10_mode1_fwd": r"dft_tile_kernelILi10ELi1ELb0ELin1ELb0ELi0E",
    "dft_tile_10_mode0_inv": r"dft_tile_kernelILi10ELi0ELb1ELi10ELb0ELi0E",
    "dft_tile_10_mode1_inv": r"dft_tile_kernelILi10ELi1ELb1ELin1ELb0ELi0E",
    "quotient_kernel_3": r"quotient_kernelILi3E",
    "aux_rows_kernel": r"aux_rows_kernel",
    "compress_kernel": r"compress_kernel",
    "quotient_kernel_full_3x256": r"quotient_kernel_fullILi3ELi256E",
    "aux_rows_kernel_full": r"aux_rows_kernel_full",
    "trace_expand_full_kernel": r"trace_expand_full_kernel",
    "trace_expand_wl_full_kernel": r"trace_expand_wl_full_kernel",
}
MUL_PIPE = ("IMAD", "IMUL")          # fmaheavy (integer multiply-add) pipe; IMAD.IADD / IMAD.MOV are adds / moves issued there
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def main():
    so = os.path.join(ROOT, "zkir_b200", "libzkir_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    bid = build_id()
    for name, pat in KERNELS.items():
        body = next((f for f in funcs[1:] if re.search(pat, f.split("\n", 1)[0])), None)
        if body is None:
            print("not found:", name)
            continue
        mangled = body.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
        hist = collections.Counter(ins)
        base = collections.Counter(i.split(".")[0] for i in ins)
        mul = sum(v for k, v in hist.items() if k.split(".")[0] in MUL_PIPE)
        lines = [f"# {TAG} SASS excerpt: {name}", f"# function {mangled}", f"# build {bid}; cuobjdump -sass zkir_b200/libzkir_b200.so (sm_100a)",
                 f"# {len(ins)} instructions; integer-multiply pipe (IMAD*/IMUL*) {mul} = {100.0 * mul / max(1, len(ins)):.1f} %",
                 "# TMA / bulk-copy mnemonics (UTMALDG, UBLKCP, LDGSTS): " + str({k: v for k, v in hist.items() if k.startswith(("UTMA", "UBLKCP", "LDGSTS"))} or "none"),
                 "", "## mnemonic histogram (full mnemonic, count)"]
        lines += [f"{k:28s} {v}" for k, v in hist.most_common(40)]
        lines += ["", "## by base mnemonic"] + [f"{k:12s} {v}" for k, v in base.most_common(25)]
        lines += ["", "## first 80 instructions"] + [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)][:80]
        with open(os.path.join(ROOT, "profiles", f"{TAG}_sass_{name}.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print(name, len(ins), "instr,", mul, "on the multiply pipe")


if __name__ == "__main__":
    main()
