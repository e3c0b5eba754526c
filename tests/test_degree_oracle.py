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
