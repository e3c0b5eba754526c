"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

    python oracle/gen_golden.py            # needs /root/reference (read-only)

The reference's `forward` is pure NumPy, but its modules import `fable`, `qiskit`
and `qiskit_aer` at the top (ChebyshevStep.py:3-5, MulStep.py:3-5, LCUStep.py:4-6,
SUMStep.py:5-7); none is installed and `forward` touches none, so empty stub
modules satisfy the imports.  Nothing from the reference is
copied: only inputs and the numbers it returns are stored.  /root/reference does
not exist on the GPU box, so the tests read the committed fixtures only.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = os.environ.get("QKAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "..", "tests", "golden")


def import_reference():
    for name, attrs in {"fable": ["fable"],
                        "qiskit": ["QuantumCircuit", "QuantumRegister", "ClassicalRegister", "transpile"],
                        "qiskit_aer": ["Aer", "AerSimulator"]}.items():
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, None)
        sys.modules.setdefault(name, mod)
    sys.path[:0] = [REF, os.path.join(REF, "QKAN_Steps_original")]
    from QKAN_Steps_original.QKANLayer import QKANLayer  # noqa
    from QKAN_Steps_original.ChebyshevStep import ChebyshevStep  # noqa
    from QKAN_Steps_original.MulStep import MulStep  # noqa
    return QKANLayer, ChebyshevStep, MulStep


def main():
    QKANLayer, ChebyshevStep, MulStep = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    # --- KAT-1..3: the reference test's own seed and draw order (QKANLayer.py:139-153)
    np.random.seed(42)
    x = np.random.uniform(-1, 1, 4)
    weights = [np.random.uniform(-1, 1, 16) for _ in range(4)]
    layer = QKANLayer(4, 4, 3)
    out1 = layer.forward(x, weights)
    out2 = layer.forward(np.zeros(4), weights)
    xb = np.array([-1.0, -1.0, 1.0, 1.0])
    out3 = layer.forward(xb, weights)
    with contextlib.redirect_stdout(io.StringIO()):
        inter = layer.get_intermediate_matrices(x, weights)
        outv = layer.forward(x, weights, verbose=True)
    np.savez(os.path.join(GOLDEN, "kat_layer_4_4_3.npz"),
             x=x, W=np.array(weights), out=out1, x_zero=np.zeros(4), out_zero=out2,
             x_boundary=xb, out_boundary=out3, out_verbose=outv,
             cheb_diag=np.array([np.diag(inter["cheb"][d]) for d in range(4)]),
             weighted_diag=np.array([np.diag(inter["weighted"][d]) for d in range(4)]),
             lcu_diag=np.diag(inter["lcu"]), reshaped=inter["reshaped"], final=inter["final"])

    # --- random batches on the shapes the reference tests and BASELINE.json name
    shapes = [(4, 4, 3, 64), (4, 8, 2, 32), (8, 4, 2, 32), (3, 2, 4, 32), (8, 8, 5, 32),
              (5, 3, 1, 32), (16, 16, 8, 8), (8, 8, 1, 16), (8, 8, 16, 16), (4, 4, 10, 16),
              (1, 1, 0, 8), (2, 2, 1, 8), (784, 10, 5, 1)]
    for (N, K, D, B) in shapes:
        rng = np.random.default_rng(1000 * N + 10 * K + D)
        X = rng.uniform(-1, 1, (B, N))
        if B > 4:
            X[1] = 0.0                       # zero input (QKANLayer.py:219)
            X[2, : N // 2] = -1.0            # boundary input (QKANLayer.py:223)
            X[2, N // 2:] = 1.0
            X[3] = 0.5                       # uniform input (QKANLayer.py:227)
        W = rng.uniform(-1, 1, (D + 1, N * K))
        lay = QKANLayer(N, K, D)
        out = np.stack([lay.forward(X[s], list(W)) for s in range(B)])
        np.savez(os.path.join(GOLDEN, f"batch_{N}_{K}_{D}.npz"), x=X, W=W, out=out)
        print(f"N={N} K={K} D={D} B={B}  max|out|={np.abs(out).max():.4f}")

    # --- out-of-range inputs: warning + clip, no error (ChebyshevStep.py:46-52)
    rng = np.random.default_rng(7)
    X = rng.uniform(-1.5, 1.5, (16, 4))
    W = rng.uniform(-1, 1, (4, 16))
    lay = QKANLayer(4, 4, 3)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = np.stack([lay.forward(X[s], list(W)) for s in range(16)])
    np.savez(os.path.join(GOLDEN, "clip_4_4_3.npz"), x=X, W=W, out=out,
             n_warnings=np.array(buf.getvalue().count("Values outside")))

    # --- step-level known answers (ChebyshevStep.py:69-102, MulStep.py:186-225)
    c2 = ChebyshevStep(2)
    t2 = c2.transform_diagonal(np.array([0.5, -0.5, 0.0]))
    dil = ChebyshevStep(1).create_dilated_chebyshev(np.array([0.5, -0.5]), 2)
    ms = MulStep(1, 4)
    ms.set_weights(1, np.array([1, .5, -.5, -1]))
    m1 = ms.get_weighted_polynomial_matrix(np.array([.5, -.5]), 2, 1)
    ms2 = MulStep(2, 4)
    ms2.set_weights(2, np.array([.5, .5, -.5, -.5]))
    m2 = ms2.get_weighted_polynomial_matrix(np.array([.5, -.5]), 2, 2)
    np.savez(os.path.join(GOLDEN, "kat_steps.npz"), t2=t2, dil=dil, mul_deg1=m1, mul_deg2=m2)
    degree_goldens(QKANLayer)
    print("golden vectors written to", os.path.normpath(GOLDEN))


class _Frame:
    """What DegreeOptimizer needs of a polars DataFrame: to_numpy() and a schema whose str() is a cache key."""

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)
        self.schema = {f"feature_{i:02d}": "Float64" for i in range(self.a.shape[1])}

    def to_numpy(self):
        return self.a


def import_reference_degree_optimizer():
    """The unmodified original_degree_optimizer/DegreeOptimizer.py.  Its top-level imports of polars, pyqubo,
    cpp_pyqubo and neal (DegreeOptimizer.py:3-6, BaseOptimizer.py:4) are not installed here and are only used by the
    QUBO search (optimize_layer) and the cross-validation helpers, which the fixtures do not call: empty stubs."""
    for name, attrs in {"polars": ["DataFrame", "col"], "cpp_pyqubo": ["Constraint"], "pyqubo": ["Array"],
                        "neal": ["SimulatedAnnealingSampler"]}.items():
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, None)
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, os.path.join(REF, "original_degree_optimizer"))
    from DegreeOptimizer import DegreeOptimizer  # noqa
    return DegreeOptimizer


def degree_goldens(QKANLayer):
    """DegreeOptimizer.evaluate_degree / is_degree_definitive / fit's weight vectors / predict (row by row)."""
    DegreeOptimizer = import_reference_degree_optimizer()
    cases = [  # (name, n, F, D, weighted, noise)
        ("a", 600, 5, 3, False, 0.05), ("b", 600, 5, 3, True, 0.05), ("c", 2500, 12, 4, True, 0.3),
        ("d", 300, 79, 2, False, 0.1), ("e", 1000, 3, 8, True, 0.01)]
    for name, n, F, D, weighted, noise in cases:
        rng = np.random.default_rng(100 + len(name) + n + F + D)
        x = rng.normal(0.0, 0.6, (n, F))                      # z-score like: a few percent beyond [-1, 1] get clipped
        y = np.cos(2.0 * x[:, 0]) + 0.3 * x[:, 1 % F] ** 3 - 0.2 * x[:, 2 % F] + noise * rng.normal(size=n)
        w = rng.uniform(0.5, 2.0, n) if weighted else None
        opt = DegreeOptimizer([F, 2], D)
        with contextlib.redirect_stdout(io.StringIO()):
            scores, comp_r2 = opt.evaluate_degree(_Frame(x), y, w)
            definitive, best = opt.is_degree_definitive(scores)
        np.savez(os.path.join(GOLDEN, f"degree_eval_{name}.npz"), x=x, y=y, w=(w if weighted else np.zeros(0)),
                 D=np.array(D), scores=scores, comp_r2=comp_r2, definitive=np.array(definitive), best=np.array(best),
                 significance_threshold=np.array(opt.significance_threshold))
        print(f"degree_eval_{name}: n={n} F={F} D={D} weighted={weighted} scores={scores}")
    # fit's weight construction (DegreeOptimizer.py:63-76) and predict (:78-95), the latter row by row because the
    # reference's own 2-D call raises (MulStep.py:62-66)
    rng = np.random.default_rng(5)
    N, K, D = 6, 3, 3
    degrees = [[int(v) for v in rng.integers(0, D + 1, N)] for _ in range(K)]
    xs = rng.normal(0.3, 1.5, (40, N))
    opt = DegreeOptimizer([N, K], D)
    opt.optimal_degrees = degrees
    opt.feature_means = np.mean(xs, axis=0)
    opt.feature_stds = np.std(xs, axis=0) + 1e-8
    opt.qkan_layer = QKANLayer(N=N, K=K, max_degree=D)
    for d in range(D + 1):                                     # the loop of fit, DegreeOptimizer.py:63-76
        wv = np.zeros(N * K)
        for out_idx, connections in enumerate(degrees):
            for in_idx, degree in enumerate(connections):
                if degree == d:
                    wv[out_idx * N + in_idx] = 1.0
        opt.qkan_layer.mul_step.set_weights(d, wv)
    W = np.array([opt.qkan_layer.mul_step._weights[d] for d in range(D + 1)])
    z = (xs - opt.feature_means) / opt.feature_stds
    with contextlib.redirect_stdout(io.StringIO()):
        pred = np.stack([opt.qkan_layer.forward(z[i], [W[d] for d in range(D + 1)]) for i in range(len(xs))])
    np.savez(os.path.join(GOLDEN, "degree_predict.npz"), x=xs, degrees=np.array(degrees), W=W, means=opt.feature_means,
             stds=opt.feature_stds, pred=pred, shape=np.array([N, K, D]))
    print("degree_predict: pred range", pred.min(), pred.max())


if __name__ == "__main__":
    main()
