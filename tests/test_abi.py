"""The C-ABI library loads on a CPU-only box and exports every symbol include/zkir_b200.h declares; proving entry
points fail loudly (no silent CPU fallback) when there is no GPU.  No compute calls here."""
import os
import re
import ctypes as C

import numpy as np
import pytest

import zkir_b200
from zkir_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "zkir_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkir_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_ffi.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zkir_b200.h but not exported"
    assert set(names) == set(_ffi.SYMBOLS), "ctypes table and header disagree"


def test_header_profile_widths_match_the_generated_air():
    text = open(os.path.join(ROOT, "include", "zkir_b200.h")).read()
    assert int(re.search(r"#define ZKIR_AIR_V2_WIDTH (\d+)u", text).group(1)) == zkir_b200.air_layout.WIDTH
    assert int(re.search(r"#define ZKIR_AIR_FULL_WIDTH (\d+)u", text).group(1)) == zkir_b200.air_layout_full.WIDTH == zkir_b200.runtime.FULL_WIDTH
    assert zkir_b200.air_layout_full.COLUMNS[:zkir_b200.air_layout.WIDTH] == zkir_b200.air_layout.COLUMNS   # the full table extends the core table


def test_proof_size_is_shape_only():
    cfg = zkir_b200.ProverConfig()
    p = cfg.params()
    s8, s20 = _ffi.lib().zkir_b200_proof_size(C.byref(p), 10), _ffi.lib().zkir_b200_proof_size(C.byref(p), 20)
    assert s8 % 4 == 0 and s20 > s8
    assert s20 == 653448   # 2^20 rows, 100 queries, 88 main + 16 aux columns, 4-row leaves, 6 fold-by-8 + 1 fold-by-4 FRI rounds: the size bench.py reports as d2h_bytes_per_step


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zkir_b200.RuntimeError) as e:
        zkir_b200.Context(0)
    assert e.value.code == _ffi.ERR_CUDA
    with pytest.raises(zkir_b200.RuntimeError):
        from conftest import fib_program
        zkir_b200.prove(fib_program(10))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", "/nonexistent/libzkir_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.lib()
