"""GPU: the degree-evaluation path through the C ABI and the drop-in DegreeOptimizer,
against the fixtures of the unmodified reference and the CPU oracle.  Tolerances: the Gram matrix and the features
are plain FP64 sums (1e-12 relative); MSE / R^2 go through normal equations + one refinement step instead of the
reference's SVD-based lstsq, bar 1e-8 relative on MSE (well-conditioned fixtures reach ~1e-13)."""
import contextlib
import glob
import io
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import degree_oracle as do

pytestmark = pytest.mark.gpu
EVAL = sorted(glob.glob(os.path.join(GOLDEN, "degree_eval_*.npz")))


@pytest.fixture(scope="module")
def Q():
    import qkan_implementation_b200 as q
    assert torch.cuda.is_available()
    return q


@pytest.mark.parametrize("path", EVAL)
def test_evaluate_degree_matches_reference(Q, path):
    g = np.load(path)
    D = int(g["D"])
    w = g["w"] if g["w"].size else None
    opt = Q.DegreeOptimizer([g["x"].shape[1], 2], D)
    with contextlib.redirect_stdout(io.StringIO()):
        scores, r2 = opt.evaluate_degree(g["x"], g["y"], w)
    assert np.abs(scores - g["scores"]).max() <= 1e-8 * np.abs(g["scores"]).max(), (scores, g["scores"])
    assert np.abs(r2 - g["comp_r2"]).max() <= 1e-7 * max(1.0, np.abs(g["comp_r2"]).max()), (r2, g["comp_r2"])
    assert opt.is_degree_definitive(scores) == (bool(g["definitive"]), int(g["best"]))


@pytest.mark.parametrize("n,F,D", [(1, 1, 0), (7, 3, 1), (33, 5, 3), (1000, 79, 3), (4097, 16, 4), (300, 2, 16), (129, 64, 0), (5000, 21, 2)])
def test_gram_and_features_match_numpy(Q, n, F, D):
    from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares
    rng = np.random.default_rng(n + F + D)
    x = rng.normal(0, 0.7, (n, F))
    y = rng.normal(size=n)
    eng = ChebyshevLeastSquares(D)
    tr = do.chebyshev_transforms(x, D)
    feats = eng.features(x)
    assert feats.shape == (D + 1, n, F)
    assert np.abs(feats - np.stack([tr[d] for d in range(D + 1)])).max() <= 1e-13 * (D + 1) ** 2
    A = np.hstack([tr[d] for d in range(D + 1)] + [y[:, None]])
    ref = A.T @ A
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    G = eng.gram(xd, yd).cpu().numpy()
    assert G.shape == ref.shape
    assert np.abs(G - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * (D + 1) ** 2
    assert np.array_equal(G, G.T)
    assert np.array_equal(G, eng.gram(xd, yd).cpu().numpy())            # deterministic


def test_predict_and_fit(Q):
    g = np.load(f"{GOLDEN}/degree_predict.npz")
    N, K, D = (int(v) for v in g["shape"])
    opt = Q.DegreeOptimizer([N, K], D)
    opt.optimal_degrees = g["degrees"].tolist()
    opt.feature_means, opt.feature_stds = g["means"], g["stds"]
    opt._build_layer()
    assert np.array_equal(opt.qkan_layer.mul_step._weights, g["W"])
    with contextlib.redirect_stdout(io.StringIO()):
        pred = opt.predict(g["x"])                                        # one launch for the whole batch
    assert pred.shape == g["pred"].shape
    assert np.abs(pred - g["pred"]).max() <= 1e-13
    # fit end to end: degrees = the QUBO's ground state for the GPU scores (oracle/degree_oracle.qubo_ground_state)
    e = np.load(EVAL[0])
    F = e["x"].shape[1]
    opt2 = Q.DegreeOptimizer([F, 2], int(e["D"]))
    with contextlib.redirect_stdout(io.StringIO()):
        opt2.fit(e["x"], e["y"])
    want = do.qubo_ground_state(e["scores"], 2 * F, opt2.complexity_weight, opt2.significance_threshold)
    assert [d for row in opt2.optimal_degrees for d in row] == want
    with contextlib.redirect_stdout(io.StringIO()):
        p2 = opt2.predict(e["x"][:50])
    ref = do.predict(e["x"][:50], opt2.feature_means, opt2.feature_stds,
                     do.fit_weight_vectors(opt2.optimal_degrees, F, 2, int(e["D"])), F, 2, int(e["D"]))
    assert np.abs(p2 - ref).max() <= 1e-13


def test_large_fit_satisfies_normal_equations(Q):
    """Size-independent property at the reference's workload shape (79 features, degree 3; 200 k of its 774 k rows):
    the residual of every fit is orthogonal to the fit's columns, and the scores decrease with the degree."""
    from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares
    n, F, D = 200_000, 79, 3
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((n, F), dtype=torch.float64, device="cuda", generator=gen) * 0.6
    y = torch.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    eng = ChebyshevLeastSquares(D)
    scores, r2 = eng.solve(x, y)
    assert np.all(np.diff(scores) <= 1e-12)
    _, t, xr = eng.residual_sums(x, y.contiguous(), None, eng.last["coef"], eng.last["ybar"], True)
    for d in range(D + 1):
        Pd = F * (d + 1)
        scale = np.sqrt(np.diag(eng.last["gram"])[:Pd] * float(t[0] + n * eng.last["ybar"] ** 2))
        assert np.abs(xr[d, :Pd] / scale).max() <= 1e-9


RES_SHAPES = [(1, 1, 0), (7, 3, 1), (65, 5, 3), (1000, 79, 3), (4097, 16, 4), (5000, 21, 2), (777, 128, 2), (513, 33, 1),
              (129, 64, 0), (64, 79, 3), (20_001, 79, 3),            # tile kernel (D <= 4, F <= 128)
              (300, 2, 16), (400, 79, 5), (257, 130, 2)]            # warp-per-sample kernel


@pytest.mark.parametrize("kernel", ["tile", "warp"])
@pytest.mark.parametrize("n,F,D", RES_SHAPES)
def test_residual_sums_match_numpy(Q, monkeypatch, kernel, n, F, D):
    """qkan_cheb_residuals against NumPy on the same inputs: residual sums of all D + 1 fits, the totals behind R^2 and
    X_D^T r_d (DegreeOptimizer.py:148-153, :277-312).  Both kernels (the tile kernel serves D <= 4, F <= 128; the
    warp-per-sample kernel everything else, and every shape when QKAN_RES_KERNEL=warp), inputs beyond [-1, 1] (clipped),
    ragged last tiles, weighted and unweighted, and an x that is not 16-byte aligned (no TMA: plain loads)."""
    from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares
    if kernel == "warp":
        monkeypatch.setenv("QKAN_RES_KERNEL", "warp")
    else:
        monkeypatch.delenv("QKAN_RES_KERNEL", raising=False)
    rng = np.random.default_rng(1000 * n + 10 * F + D)
    D1, P = D + 1, F * (D + 1)
    xh = rng.normal(0.0, 0.7, (n, F))
    yh = rng.normal(0.3, 1.0, n)
    wh = rng.uniform(0.5, 1.5, n)
    coef = np.zeros((D1, P))
    for d in range(D1):
        coef[d, :F * (d + 1)] = rng.normal(0.0, 0.3, F * (d + 1))
    ybar = float(yh.mean())
    xc = np.clip(xh, -1.0, 1.0)
    T = [np.ones_like(xc), xc]
    for k in range(2, D1):
        T.append(2.0 * xc * T[-1] - T[-2])
    X = np.hstack(T[:D1])                                    # degree-major columns k F + f, the reference's np.hstack order
    R = yh[:, None] - X @ coef.T                             # [n, D+1]
    eng = ChebyshevLeastSquares(D)
    for aligned in (True, False):
        if aligned:
            xd = torch.from_numpy(xh).cuda()
        else:
            base = torch.empty(n * F + 1, dtype=torch.float64, device="cuda")
            xd = base[1:].view(n, F)
            xd.copy_(torch.from_numpy(xh))
            assert xd.data_ptr() % 16 == 8
        yd = torch.from_numpy(yh).cuda()
        for wd, wref in ((torch.from_numpy(wh).cuda(), wh), (None, np.ones(n))):
            for want_xtr in (True, False):
                s, t, xr = eng.residual_sums(xd, yd, wd, coef, ybar, want_xtr)
                ref_s = np.stack([(R ** 2).sum(0), (wref[:, None] * R ** 2).sum(0)], axis=1)
                ref_t = np.array([((yh - ybar) ** 2).sum(), (wref * yh * yh).sum(), wref.sum(), yh.sum()])
                assert np.abs(s - ref_s).max() <= 1e-11 * np.abs(ref_s).max(), (kernel, aligned, want_xtr)
                assert np.abs(t - ref_t).max() <= 1e-11 * np.abs(ref_t).max()
                if want_xtr:
                    ref_x = (X.T @ R).T                      # [D+1, P]
                    assert np.abs(xr - ref_x).max() <= 1e-11 * max(1.0, np.abs(ref_x).max()), (kernel, aligned)
