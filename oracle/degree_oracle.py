"""CPU oracle for the degree-evaluation path.  TEST INFRASTRUCTURE ONLY: imported by
tests/, oracle/gen_golden.py and the CPU-baseline leg of tools/bench_degree.py, never by the product path.

NumPy restatement of original_degree_optimizer/DegreeOptimizer.py of the reference (file:line cited per
function).  Parity pinned: tests/golden/degree_*.npz hold the outputs of the UNMODIFIED reference class
(imported by oracle/gen_golden.py with polars / pyqubo / neal stubbed - evaluate_degree, _compute_metrics,
is_degree_definitive and the weight construction of fit touch none of them), and tests/test_oracle.py checks
this restatement against them.
"""
import numpy as np

from . import qkan_oracle


def chebyshev_transforms(feature_data: np.ndarray, max_degree: int) -> dict:
    """DegreeOptimizer._compute_transforms (DegreeOptimizer.py:96-119): transforms[d] = T_d of every feature
    column, through ChebyshevStep.transform_diagonal (ChebyshevStep.py:32-53: clip to [-1, 1], cos(d arccos x))."""
    x = np.clip(np.asarray(feature_data, dtype=np.float64), -1.0, 1.0)
    return {d: np.cos(d * np.arccos(x)) for d in range(max_degree + 1)}


def compute_metrics(y_true, y_pred, weights=None) -> dict:
    """DegreeOptimizer._compute_metrics (DegreeOptimizer.py:277-312), kept as written: the weighted branch calls
    sum(w err^2) 'ss_tot' and sum(w y^2) 'ss_res', and r2 = 1 - ss_tot / ss_res in both branches."""
    y_true = np.asarray(y_true, dtype=np.float64).reshape(-1, 1)
    y_pred = np.asarray(y_pred, dtype=np.float64).reshape(-1, 1)
    sq = (y_true - y_pred) ** 2
    if weights is not None:
        w = np.asarray(weights, dtype=np.float64).reshape(-1, 1)
        mse = np.average(sq, weights=w)
        ss_tot = np.sum(w * sq)
        ss_res = np.sum(w * y_true ** 2)
    else:
        mse = np.mean(sq)
        ss_tot = np.sum((y_true - np.mean(y_true)) ** 2)
        ss_res = np.sum(sq)
    r2 = 0.0 if ss_tot < np.finfo(float).eps else 1 - ss_tot / ss_res
    return {"mse": float(mse), "r2": float(r2)}


def evaluate_degree(feature_data, y_data, max_degree: int, weights=None):
    """DegreeOptimizer.evaluate_degree (DegreeOptimizer.py:122-158): for d = 0..D least squares of y on
    [T_0 | ... | T_d] (np.linalg.lstsq, minimum norm), scores[d] = MSE, comp_r2[d] = the reference's R^2."""
    tr = chebyshev_transforms(feature_data, max_degree)
    y = np.asarray(y_data, dtype=np.float64)
    scores = np.zeros(max_degree + 1)
    comp_r2 = np.zeros(max_degree + 1)
    for d in range(max_degree + 1):
        X = np.hstack([tr[k].reshape(len(y), -1) for k in range(d + 1)])
        coeffs = np.linalg.lstsq(X, y, rcond=None)[0]
        m = compute_metrics(y, X @ coeffs, weights)
        scores[d], comp_r2[d] = m["mse"], m["r2"]
    return scores, comp_r2


def is_degree_definitive(scores, significance_threshold: float):
    """DegreeOptimizer.is_degree_definitive (DegreeOptimizer.py:159-181)."""
    best = int(np.argmin(scores))
    best_score = float(scores[best])
    for d in range(len(scores)):
        if d != best:
            score = float(scores[d])
            if (score - best_score) / (score + 1e-10) < significance_threshold:
                return False, best
    return True, best


def qubo_ground_state(scores, num_functions: int, complexity_weight: float, significance_threshold: float):
    """The minimiser of the QUBO that DegreeOptimizer.optimize_layer (DegreeOptimizer.py:183-253) hands to the
    annealer.  The objective is a sum over functions i of  sum_d a_d q[i, d] + 10 (sum_d q[i, d] - 1)^2  with the same
    a_d for every i, so its ground state is one-hot per function at argmin_d a_d (a_d < 10 in every case that
    matters; ties go to the lowest degree):
      definitive degree d*:  a_d = -100 at d*, +100 elsewhere                                  (:214-219)
      otherwise:             a_d = -(scores[d] - scores[d-1]) + cw d^2   (d = 0: -scores[0])   (:221-225)"""
    definitive, best = is_degree_definitive(scores, significance_threshold)
    if definitive:
        return [best] * num_functions
    D1 = len(scores)
    a = np.array([-(scores[d] - scores[d - 1] if d > 0 else scores[d]) + complexity_weight * d * d for d in range(D1)])
    return [int(np.argmin(a))] * num_functions


def fit_weight_vectors(optimal_degrees, N: int, K: int, max_degree: int):
    """DegreeOptimizer.fit (DegreeOptimizer.py:63-76): weights[d][out_idx * N + in_idx] = 1 where the connection's
    degree is d."""
    W = np.zeros((max_degree + 1, N * K))
    for out_idx, connections in enumerate(optimal_degrees):
        for in_idx, degree in enumerate(connections):
            W[degree, out_idx * N + in_idx] = 1.0
    return W


def predict(feature_data, feature_means, feature_stds, W, N: int, K: int, max_degree: int):
    """DegreeOptimizer.predict (DegreeOptimizer.py:78-95) row by row (the reference's own 2-D call raises in
    MulStep.py:62-66): z-score, then QKANLayer.forward (which clips, ChebyshevStep.py:52)."""
    z = (np.asarray(feature_data, dtype=np.float64) - feature_means) / feature_stds
    return qkan_oracle.forward_closed_form(z, W, N, K, max_degree)
