import os
import sys
import ctypes as C

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P = 2013265921


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


from oracle.binding import Oracle  # noqa: E402,F401


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def gpu_ctx():
    import zkir_b200
    ctx = zkir_b200.Context(0)   # raises loudly when the .so or the GPU is missing
    yield ctx
    ctx.close()


from zkir_b200.workloads import FIB_SRC, ADD_SRC, fib_program, fib_program_input, fib_trace  # noqa: E402,F401
