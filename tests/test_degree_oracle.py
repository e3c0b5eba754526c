"""CPU: the degree-evaluation oracle (oracle/degree_oracle.py) against fixtures produced by the UNMODIFIED reference
DegreeOptimizer (oracle/gen_golden.py: degree_goldens), and the host logic of the drop-in DegreeOptimizer that needs
no GPU."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import degree_oracle as do

EVAL = sorted(glob.glob(os.path.join(GOLDEN, "degree_eval_*.npz")))


@pytest.mark.parametrize("path", EVAL)
def test_evaluate_degree_oracle_matches_reference(path):
    g = np.load(path)
    w = g["w"] if g["w"].size else None
    scores, r2 = do.evaluate_degree(g["x"], g["y"], int(g["D"]), w)
    assert np.abs(scores - g["scores"]).max() <= 1e-12 * np.abs(g["scores"]).max()
    assert np.abs(r2 - g["comp_r2"]).max() <= 1e-9 * max(1.0, np.abs(g["comp_r2"]).max())
    definitive, best = do.is_degree_definitive(scores, float(g["significance_threshold"]))
    assert definitive == bool(g["definitive"]) and best == int(g["best"])


def test_fit_weights_and_predict_oracle_match_reference():
    g = np.load(f"{GOLDEN}/degree_predict.npz")
    N, K, D = (int(v) for v in g["shape"])
    W = do.fit_weight_vectors(g["degrees"].tolist(), N, K, D)
    assert np.array_equal(W, g["W"])
    pred = do.predict(g["x"], g["means"], g["stds"], W, N, K, D)
    assert np.abs(pred - g["pred"]).max() <= 4e-16


def test_qubo_ground_state():
    # definitive: one degree clearly best -> every function takes it (DegreeOptimizer.py:214-219)
    assert do.qubo_ground_state(np.array([0.5, 0.4, 0.01, 0.3]), 6, 0.1, 0.05) == [2] * 6
    # not definitive: argmin of -(improvement) + cw d^2 (:221-225)
    s = np.array([0.40, 0.39, 0.385, 0.384])
    a = [-(s[0]), -(s[1] - s[0]) + 0.1, -(s[2] - s[1]) + 0.4, -(s[3] - s[2]) + 0.9]
    assert do.qubo_ground_state(s, 4, 0.1, 0.05) == [int(np.argmin(a))] * 4


def test_host_logic_without_gpu(tmp_path):
    from qkan_implementation_b200 import DegreeOptimizer
    opt = DegreeOptimizer([5, 2], 3)
    assert (opt.num_layers, opt.max_degree, opt.complexity_weight, opt.significance_threshold) == (1, 3, 0.1, 0.05)
    for path in EVAL:
        g = np.load(path)
        opt.significance_threshold = float(g["significance_threshold"])
        assert opt.is_degree_definitive(g["scores"]) == (bool(g["definitive"]), int(g["best"]))
    rng = np.random.default_rng(0)
    y, p, w = rng.normal(size=50), rng.normal(size=50), rng.uniform(0.5, 2, 50)
    for ww in (None, w):
        m, ref = opt._compute_metrics(y, p, ww), do.compute_metrics(y, p, ww)
        assert m == ref
    with pytest.raises(RuntimeError, match="Not fitted yet"):
        opt.predict(np.zeros((3, 5)))
    f = str(tmp_path / "state.npy")
    opt.save_state(f)
    other = DegreeOptimizer([1, 1], 1)
    other.load_state(f, {'n_rows': 1, 'columns': [], 'sort_by': 'x'})
    assert other.network_shape == [5, 2] and other.max_degree == 3 and other.data_same is False


def _design(x, D):
    xc = np.clip(x, -1.0, 1.0)
    T = [np.ones_like(xc), xc]
    for _ in range(2, D + 1):
        T.append(2.0 * xc * T[-1] - T[-2])
    return np.hstack(T[:D + 1])


@pytest.mark.parametrize("F,D", [(1, 0), (3, 1), (5, 3), (79, 3), (16, 4), (2, 8)])
def test_nested_solver_returns_lstsq_minimum_norm_solutions(F, D):
    """Host side of evaluate_degree (no GPU): from ONE Gram matrix of [X_D | y], _NestedSolver returns, for every degree d, what
    np.linalg.lstsq(X_d, y) returns in the reference (DegreeOptimizer.py:146) - the minimum-norm solution of the rank-deficient
    system (the T_0 columns of all features coincide), through one Cholesky factor of the reduced matrix."""
    from qkan_implementation_b200.degree_optimizer import _NestedSolver, _blas_single_thread
    rng = np.random.default_rng(100 * F + D)
    n = 4000
    x = rng.normal(0.0, 0.6, (n, F))
    y = np.cos(2 * x[:, 0]) + 0.1 * rng.normal(size=n)
    X = _design(x, D)
    A = np.hstack([X, y[:, None]])
    G = A.T @ A
    P = F * (D + 1)
    with _blas_single_thread():
        solver = _NestedSolver(G, n, F, D)
        assert solver.L is not None                          # well conditioned: the Cholesky path
        for d in range(D + 1):
            Pd = F * (d + 1)
            got = solver.solve(d, G[:Pd, P])
            ref = np.linalg.lstsq(X[:, :Pd], y, rcond=None)[0]
            assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (d, np.abs(got - ref).max())


def test_nested_solver_falls_back_on_duplicated_features():
    """A duplicated feature makes the reduced Gram matrix singular: the solver must leave the Cholesky path and still return
    lstsq's minimum-norm solution (eigen-decomposition of each leading block with lstsq's rank cut-off)."""
    from qkan_implementation_b200.degree_optimizer import _NestedSolver
    rng = np.random.default_rng(7)
    n, F, D = 3000, 4, 2
    x = rng.normal(0.0, 0.6, (n, F))
    x[:, 3] = x[:, 1]
    y = x[:, 0] ** 2 + 0.05 * rng.normal(size=n)
    X = _design(x, D)
    A = np.hstack([X, y[:, None]])
    G = A.T @ A
    P = F * (D + 1)
    solver = _NestedSolver(G, n, F, D)
    assert solver.L is None
    for d in range(D + 1):
        Pd = F * (d + 1)
        got = solver.solve(d, G[:Pd, P])
        ref = np.linalg.lstsq(X[:, :Pd], y, rcond=None)[0]
        assert np.abs(X[:, :Pd] @ (got - ref)).max() <= 1e-7 * np.abs(y).max()      # the same fit
        assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
