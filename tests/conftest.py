import glob
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def ensure_library_built():
    """The CUDA library is a build artefact (git-ignored): build it once if the tree is fresh.
    nvcc cross-compiles sm_100a without a GPU (about a minute on 8 cores)."""
    lib = os.path.join(ROOT, "qkan_implementation_b200", "libqkan_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()


def golden_batches():
    """[(N, K, D, path)] of the fixtures written by oracle/gen_golden.py from the unmodified reference."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "batch_*.npz"))):
        N, K, D = map(int, re.findall(r"\d+", os.path.basename(f)))
        out.append((N, K, D, f))
    return out


def rel_err(a, b):
    """per-sample relative 2-norm error (the reference tests use relative Frobenius norms)."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    b = np.atleast_2d(np.asarray(b, dtype=np.float64))
    num = np.linalg.norm(a - b, axis=1)
    den = np.maximum(np.linalg.norm(b, axis=1), 1e-300)
    return float((num / den).max())


@pytest.fixture(scope="session")
def emu():
    import ctypes
    import __graft_entry__ as g
    so = g.build_emu()
    lib = ctypes.CDLL(so)
    lib.qkan_emu_forward.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib
