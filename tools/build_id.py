#!/usr/bin/env python3
"""Identity of the CUDA build: sha256 over the device-side sources of the library (zkir_b200/csrc/*.cu, *.cuh, *.h and include/*.h; the
host-only files under csrc/host/ -- interpreter, packer, verifier -- do not change any kernel), first 12 hex digits.
bench.py prints it, the profile tools store it, so a number read from profiles/ can be matched to the build that produced it."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_id():
    h = hashlib.sha256()
    files = []
    for d in ("zkir_b200/csrc", "include"):
        for f in sorted(os.listdir(os.path.join(ROOT, d))):
            if f.endswith((".cu", ".cc", ".h", ".cuh")):
                files.append(os.path.join(d, f))
    for rel in files:
        h.update(rel.encode())
        h.update(open(os.path.join(ROOT, rel), "rb").read())
    return h.hexdigest()[:12]


if __name__ == "__main__":
    print(build_id())
