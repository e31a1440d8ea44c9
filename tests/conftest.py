import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "lnlike_golden.npz"))


@pytest.fixture(scope="session")
def built():
    """The native pieces, compiled if stale (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    """The CUDA engine; GPU tests fail loudly (not skip) if the extension or the device is missing."""
    from gprotation_b200 import GPEngine

    eng = GPEngine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def oracle_c(built):
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libgp_oracle.so"))
    for f in (lib.gp_oracle_lnlike_f64, lib.gp_oracle_lnlike_ld):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    def call(which, th, t, y, e):
        th, t, y, e = (np.ascontiguousarray(a, dtype=np.float64) for a in (th, t, y, e))
        fn = lib.gp_oracle_lnlike_ld if which == "ld" else lib.gp_oracle_lnlike_f64
        return fn(th.ctypes.data, len(t), t.ctypes.data, y.ctypes.data, e.ctypes.data)

    return call
