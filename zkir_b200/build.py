"""Build recipe for libzkir_b200.so (CUDA kernels + C ABI + host-side VM/packer/verifier), sm_100a only.

nvcc cross-compiles without a GPU.  Objects are cached under zkir_b200/csrc/_obj keyed by source mtime so a
rebuild after touching one file takes seconds.  The .so is built IN-TREE so it travels with gpurun snapshots.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libzkir_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = GENCODE + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX,
                        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "--expt-relaxed-constexpr",
                        "-Xptxas", "-v"] + os.environ.get("ZKIR_NVCC_EXTRA", "").split()   # experiments, e.g. ZKIR_NVCC_EXTRA=-DZKIR_P2_PIN=0


def sources():
    out = []
    for d in (CSRC, os.path.join(CSRC, "host")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cc")):
                out.append(os.path.join(d, f))
    return out


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(CSRC, "host"), os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


# translation units that instantiate the generated AIR: compiled once per profile (csrc/air_profile.h)
PER_PROFILE = ("quotient.cu", "aux_gen.cu", "trace_expand.cu", "pack.cc", "verify.cc")


def _compile(job, hdr_m, verbose):
    src, full = job
    obj = os.path.join(OBJ, os.path.basename(src) + (".full" if full else "") + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_m):
        return obj, ""
    cmd = [NVCC] + NVCC_FLAGS + (["-DZKIR_PROFILE_FULL"] if full else []) + (["-x", "cu"] if src.endswith(".cu") else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    hdr_m = _headers_mtime()
    srcs = [(f, False) for f in sources()] + [(f, True) for f in sources() if os.path.basename(f) in PER_PROFILE]
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, hdr_m, verbose), srcs))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if verbose and log:
        print(log)
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + GENCODE + ["-shared", "-ccbin", HOST_CXX, "-o", LIB] + objs + ["-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
