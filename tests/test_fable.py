"""FABLE gate lists.  CPU: host-side construction against a dense NumPy
simulator (oracle/circuit_sim.py).  GPU: the reference's own block-encoding unit tests, with
qkan_simulate_circuit in place of Qiskit Aer (MulStep.py:186-264, LCUStep.py:109-211,
SUMStep.py:80-187, ChebyshevStep.py:117-134)."""
import numpy as np
import pytest

from oracle import circuit_sim as cs
from oracle import qkan_oracle as o
from qkan_implementation_b200.fable import fable, gray_permute, sfwht, verify_unitary


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("kind", ["dense", "diag", "zero", "big"])
def test_block_encoding_identity_cpu(dim, kind):
    rng = np.random.default_rng(dim * 10 + len(kind))
    A = rng.uniform(-1, 1, (dim, dim))
    if kind == "diag":
        A = np.diag(rng.uniform(-1, 1, dim))
    elif kind == "zero":
        A = np.zeros((dim, dim))
    elif kind == "big":
        A = A * 7.5                                   # alpha > 1 renormalisation (SUMStep.py:169-187)
    circ, alpha = fable(A, 0)
    n = circ.n
    assert circ.num_qubits == 2 * n + 1
    ops = circ.count_ops()
    assert ops["h"] == 2 * n and ops["swap"] == n
    if kind == "zero":
        assert ops.get("ry", 0) == 1 and "cx" not in ops          # collapses to a single Ry(pi)
    elif kind == "dense":
        assert ops["ry"] == 4 ** n and ops["cx"] == 4 ** n
    blk = cs.top_left_block(circ.gates, circ.params, circ.num_qubits, 1 << n)
    Ap = np.zeros((1 << n, 1 << n))
    Ap[:dim, :dim] = A
    assert np.abs(blk.real * alpha * (1 << n) - Ap).max() < 2e-14
    assert np.abs(blk.imag).max() == 0.0
    assert alpha == 1.0 if kind != "big" else alpha > 1.0


def test_angle_transform_matches_definition():
    rng = np.random.default_rng(3)
    th = rng.uniform(0, np.pi, 16)
    got = gray_permute(sfwht(th))
    k = 4
    want = np.array([sum(th[j] * (-1) ** bin(j & (i ^ (i >> 1))).count("1") for j in range(16)) / 16 for i in range(16)])
    assert np.abs(got - want).max() < 1e-15


# ------------------------------------------------------------------ GPU: the reference's unit tests
gpu = pytest.mark.gpu


@gpu
def test_gpu_simulator_matches_dense_oracle():
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, (4, 4))
    circ, alpha = fable(A, 0)
    cols = circ.columns(np.arange(1 << circ.num_qubits))
    ref = np.stack([cs.evolve(circ.gates, circ.params, circ.num_qubits, j) for j in range(1 << circ.num_qubits)])
    assert np.abs(cols - ref).max() < 1e-14
    U = cols.T
    assert np.abs(U.conj().T @ U - np.eye(U.shape[0])).max() < 1e-13       # unitary


@gpu
def test_mulstep_block_encodings():
    import qkan_implementation_b200 as Q
    x = np.array([0.5, -0.5])
    ms = Q.MulStep(1, 4)
    ms.set_weights(1, np.array([1, .5, -.5, -1]))
    circ, alpha = ms.create_weighted_chebyshev(x, 2, 1)                      # MulStep.py:186-209
    expected = np.diag([.5, .25, .25, .5])
    assert verify_unitary(circ, expected, alpha) < 1e-6
    blk = circ.block().real * alpha * 4
    assert np.array_equal(np.abs(blk) > 1e-10, np.abs(expected) > 1e-10)     # same non-zero pattern (:153-166)
    ms2 = Q.MulStep(2, 4)
    ms2.set_weights(2, np.array([.5, .5, -.5, -.5]))
    circ, alpha = ms2.create_weighted_chebyshev(x, 2, 2)                     # MulStep.py:211-234
    assert verify_unitary(circ, np.diag([-.25, -.25, .25, .25]), alpha) < 1e-6
    ms3 = Q.MulStep(1, 4)                                                     # zero weights -> zero block (:249-264)
    circ, alpha = ms3.create_weighted_chebyshev(x, 2, 1)
    assert np.abs(circ.block()).max() < 1e-12


@gpu
@pytest.mark.parametrize("N,K,d", [(4, 4, 5), (4, 8, 8), (8, 4, 7), (4, 8, 20)])
def test_lcu_block_encodings(N, K, d):
    import qkan_implementation_b200 as Q
    rng = np.random.default_rng(42)                                           # LCUStep.py:67
    x = rng.uniform(-1, 1, N)
    ms = Q.MulStep(d, N * K)
    W = rng.uniform(-1, 1, (d + 1, N * K))
    for deg in range(d + 1):
        ms.set_weights(deg, W[deg])
    lcu = Q.LCUStep(d)
    circ, alpha = lcu.combine_weighted_polynomials(x, ms, K)                 # LCUStep.py:109-161
    expected = np.diag(o.stage_diagonals(x, W, N, K, d)["lcu"][0])
    assert circ.num_qubits == 2 * int(np.ceil(np.log2(N * K))) + 1
    assert verify_unitary(circ, expected, alpha) < 1e-6


@gpu
def test_sum_block_encodings():
    import qkan_implementation_b200 as Q
    s = Q.SUMStep()
    circ, alpha = s.apply_sum(np.diag([1, .5, -.5, -1.0]), 2, 2)             # SUMStep.py:80-102
    assert verify_unitary(circ, np.diag([0.75, -0.75]), alpha) < 1e-6
    rng = np.random.default_rng(42)
    for N, K in ((4, 4), (4, 8), (8, 4)):                                     # SUMStep.py:104-130
        M = np.diag(rng.uniform(-1, 1, N * K))
        circ, alpha = s.apply_sum(M, N, K)
        expected = np.diag(np.sum(np.diag(M).reshape(N, K, order="F"), axis=0) / N)
        assert verify_unitary(circ, expected, alpha) < 1e-6
    for scale in (1e-3, 1e-2, 1.0, 10.0, 100.0):                              # SUMStep.py:169-187 (alpha > 1)
        M = np.diag(rng.uniform(-1, 1, 16)) * scale
        circ, alpha = s.apply_sum(M, 4, 4)
        expected = np.diag(np.sum(np.diag(M).reshape(4, 4, order="F"), axis=0) / 4)
        assert verify_unitary(circ, expected, alpha) < 1e-6


@gpu
def test_chebyshev_dilated_block_encoding():
    import qkan_implementation_b200 as Q
    cheb = Q.ChebyshevStep(8)                                                 # ChebyshevStep.py:117-134
    errs = []
    for seed in range(40):                                                    # the reference draws one unseeded x and asks < 1e-15
        x = np.random.default_rng(seed).uniform(-1, 1, 4)
        A = cheb.create_dilated_chebyshev(x, 1)
        circ, alpha = fable(A, 0)
        blk = circ.block().real * alpha * 4
        errs.append(np.linalg.norm(blk - A) / np.linalg.norm(A))
    # measured on B200 (tools/cheb_be_error.py): median 6.4e-16, max 1.0e-15 - at the reference's bar, a few ulp
    assert np.median(errs) < 1e-15 and max(errs) < 2e-15, (np.median(errs), max(errs))


@gpu
def test_aer_shim_full_unitary_on_gpu_matches_oracle():
    """tests/shims/qiskit_aer backed by qkan_simulate_circuit: the FULL unitary of an 11-qubit block-encoding (4x8 LCU
    matrix, the largest circuit of the reference's tests) equals the oracle's dense simulation, and it is unitary."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims"))
    try:
        import qiskit_aer
        rng = np.random.default_rng(5)
        A = np.diag(rng.uniform(-1, 1, 32))
        circ, alpha = fable(A, 0)
        assert circ.num_qubits == 11
        sim = qiskit_aer.Aer.get_backend("unitary_simulator")
        U = np.asarray(sim.run(circ).result().get_unitary(circ))
        assert qiskit_aer.BACKEND_USED[-1] == "gpu" and U.shape == (2048, 2048)
        ref = cs.unitary(circ.gates, circ.params, circ.num_qubits)
        assert np.abs(U - ref).max() < 1e-13
        assert np.abs(U.conj().T @ U - np.eye(2048)).max() < 1e-12
        assert np.abs(U[:32, :32].real * alpha * 32 - A).max() < 1e-13
    finally:
        sys.path.pop(0)
        for m in ("qiskit_aer",):
            sys.modules.pop(m, None)


@gpu
def test_c3_size_lcu_block_encoding_17_qubits():
    """The LCU block-encoding of BASELINE configs[2] (N16 K16 D8: a 256 x 256 diagonal, 17 qubits, 65 536 Ry + 65 536 CX in the
    oracle): the run-fused simulator evaluates its 256 block columns; identity U[:256, :256] alpha 256 = A as in
    LCUStep.py:69-107 (relative Frobenius error < 1e-6 there)."""
    import qkan_implementation_b200 as Q
    rng = np.random.default_rng(42)
    N = K = 16
    d = 8
    x = rng.uniform(-1, 1, N)
    ms = Q.MulStep(d, N * K)
    W = rng.uniform(-1, 1, (d + 1, N * K))
    for deg in range(d + 1):
        ms.set_weights(deg, W[deg])
    circ, alpha = Q.LCUStep(d).combine_weighted_polynomials(x, ms, K)
    assert circ.num_qubits == 17 and circ.count_ops()["ry"] == 4 ** 8
    expected = np.diag(o.stage_diagonals(x, W, N, K, d)["lcu"][0])
    assert verify_unitary(circ, expected, alpha) < 1e-12


def test_oracle_run_fused_unitary_matches_gate_by_gate():
    """oracle/circuit_sim.unitary (run fusion) against the plain gate-by-gate evolve, on a circuit that mixes every gate kind."""
    rng = np.random.default_rng(8)
    circ, _ = fable(rng.uniform(-1, 1, (4, 4)), 0)
    gates = list(circ.gates) + [(cs.X, 1, 0), (cs.Z, 1, 0), (cs.H, 1, 0), (cs.CX, 0, 1), (cs.SWAP, 0, 3), (cs.RY, 3, 0), (cs.CX, 2, 3)]
    params = list(circ.params) + [0, 0, 0, 0, 0, 0.37, 0]
    U = cs.unitary(gates, params, circ.num_qubits)
    ref = np.stack([cs.evolve(gates, params, circ.num_qubits, j) for j in range(1 << circ.num_qubits)], axis=1)
    assert np.abs(U - ref).max() < 1e-14
