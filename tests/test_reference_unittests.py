"""The reference's OWN in-file unit tests, unmodified, run from /root/reference with the third-party packages they import
replaced by tests/shims (MulStep.py:110-264, LCUStep.py:63-211, SUMStep.py:34-187,
ChebyshevStep.py:68-134).  Two legs:

* impl="reference": the reference's step classes build the matrices; `fable(...)` is the product's gate-list generator and
  `Aer ... get_unitary` the product's simulator (GPU) or the oracle's (no GPU).  This checks the product's FABLE circuits
  against the reference's own verify_unitary code and tolerances.
* impl="package" (needs a GPU as well): the reference test classes run against THIS package's drop-in step classes
  (the module globals the tests look up - MulStep, LCUStep, SUMStep, ChebyshevStep - are pointed at the package).

/root/reference does not exist on the GPU box: there these tests skip and tests/test_fable.py carries the GPU leg.
One reference test is wrong as written and fails against the reference itself (ChebyshevStep.py:111-112 expects
transform_diagonal([1.5, 0.5]) to raise although :52 clips first): it is the only expected failure."""
import importlib
import io
import os
import sys
import unittest
from contextlib import redirect_stdout

import pytest

REF = "/root/reference"
STEPS = os.path.join(REF, "QKAN_Steps_original")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
pytestmark = pytest.mark.skipif(not os.path.isdir(STEPS), reason="the reference checkout is not on this machine")

EXPECTED_FAILURES = {"test_input_validation"}        # ChebyshevStep.py:104-115, broken in the reference itself
CASES = [("ChebyshevStep", "TestChebyshevStep"), ("MulStep", "TestMulStep"), ("LCUStep", "TestLCUStep"), ("SUMStep", "TestSUMStep")]


@pytest.fixture(scope="module")
def ref_modules():
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for name in ("fable", "qiskit", "qiskit_aer", "ChebyshevStep", "MulStep", "LCUStep", "SUMStep"):
        sys.modules.pop(name, None)
    sys.path[:0] = [SHIMS, STEPS, REF]
    mods = {m: importlib.import_module(m) for m, _ in CASES}
    yield mods
    sys.path[:] = saved_path
    for name in ("fable", "qiskit", "qiskit_aer", "ChebyshevStep", "MulStep", "LCUStep", "SUMStep"):
        sys.modules.pop(name, None)
    sys.modules.update({k: v for k, v in saved_mods.items() if k in ("fable", "qiskit", "qiskit_aer")})


def _run(case_cls):
    # TestChebyshevStep.test_dilated_block_encoding_different_sizes draws an UNSEEDED x and asks for a relative error
    # < 1e-15, which is the rounding floor of a 16-angle double-precision Gray-code decomposition (this package over 40
    # draws: median 6.4e-16, max 1.06e-15, see test_block_encoding_error_distribution below and tools/cheb_be_error.py),
    # so an unseeded run fails now and then.  Seeding NumPy from outside keeps the suite deterministic without
    # touching the reference's test.
    import numpy as np
    np.random.seed(0)
    suite = unittest.defaultTestLoader.loadTestsFromTestCase(case_cls)
    buf = io.StringIO()
    with redirect_stdout(buf):                        # the reference tests print every matrix
        res = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(suite)
    bad = [(t.id().split(".")[-1], tb) for t, tb in res.failures + res.errors]
    return res.testsRun, bad


@pytest.mark.parametrize("module,cls", CASES)
def test_reference_unit_tests_with_product_fable_and_simulator(ref_modules, module, cls):
    import qiskit_aer
    qiskit_aer.BACKEND_USED.clear()
    ran, bad = _run(getattr(ref_modules[module], cls))
    unexpected = [(n, tb) for n, tb in bad if n not in EXPECTED_FAILURES]
    assert ran >= 2 and not unexpected, "\n".join(f"{n}:\n{tb}" for n, tb in unexpected)
    if module != "ChebyshevStep":
        assert qiskit_aer.BACKEND_USED, "no circuit was simulated"


@pytest.mark.gpu
@pytest.mark.parametrize("module,cls", CASES)
def test_reference_unit_tests_against_the_package_classes(ref_modules, module, cls):
    """Drop-in check: the reference's test classes, with the step classes they name resolved to this package's."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU (the package has no CPU path)")
    import qkan_implementation_b200 as Q
    mod = ref_modules[module]
    saved = {n: getattr(mod, n) for n in ("ChebyshevStep", "MulStep", "LCUStep", "SUMStep") if hasattr(mod, n)}
    try:
        for n in saved:
            setattr(mod, n, getattr(Q, n))
        ran, bad = _run(getattr(mod, cls))
    finally:
        for n, v in saved.items():
            setattr(mod, n, v)
    unexpected = [(n, tb) for n, tb in bad if n not in EXPECTED_FAILURES]
    assert ran >= 2 and not unexpected, "\n".join(f"{n}:\n{tb}" for n, tb in unexpected)


def test_block_encoding_error_distribution():
    """The quantity ChebyshevStep.py:117-134 asserts on one unseeded draw, over 40 seeded draws, with the product's FABLE
    generator and the oracle's simulator: at the reference's 1e-15 bar (a few ulp), never far above it."""
    import numpy as np
    from oracle import circuit_sim as cs
    from qkan_implementation_b200.fable import fable
    errs = []
    for seed in range(40):
        x = np.random.default_rng(seed).uniform(-1, 1, 4)
        A = np.diag(np.cos(8 * np.arccos(x)))                          # create_dilated_chebyshev(x, 1), degree 8
        circ, alpha = fable(A, 0)
        blk = cs.top_left_block(circ.gates, circ.params, circ.num_qubits, 4).real * alpha * 4
        errs.append(np.linalg.norm(blk - A) / np.linalg.norm(A))
    assert np.median(errs) < 1e-15 and max(errs) < 2e-15, (np.median(errs), max(errs))
