"""Host mirror of the reference's DegreeOptimizer (original_degree_optimizer/DegreeOptimizer.py) on the B200 path
(SURVEY 8(f) ranks 3 and 4): same constructor, attributes and method names; `predict` runs the batch through the
drop-in QKANLayer, `evaluate_degree` runs the Chebyshev-feature least squares on the GPU (csrc/qkan_degree.cu),
and the QUBO of `optimize_layer` - which is separable per function - is minimised in closed form instead of by
pyqubo + neal simulated annealing.  There is no CPU fallback: every numeric method needs libqkan_b200.so and a GPU.

Accepted data containers: anything with ``to_numpy()`` (the reference passes polars DataFrames), NumPy arrays and
torch tensors (CPU or CUDA).
"""
import ctypes
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _binding as _b
from .layer import QKANLayer


def _as_numpy(a) -> np.ndarray:
    if hasattr(a, "to_numpy"):
        a = a.to_numpy()
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def _as_device(a, device) -> torch.Tensor:
    if hasattr(a, "to_numpy"):
        a = a.to_numpy()
    if not isinstance(a, torch.Tensor):
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
    return a.to(device=device, dtype=torch.float64).contiguous()


class _NestedSolver:
    """Minimum-norm least-squares solves X_d c = rhs for every degree d from the one Gram matrix of X_D.

    The T_0 columns of all F features are the same all-ones column, so X_d always has rank <= 1 + F d and np.linalg.lstsq
    returns the minimum-norm solution, which gives each of the F copies 1 / F of the constant's coefficient.  Fast path:
    drop the copies (one ones column + the degree >= 1 columns) and take ONE Cholesky factor of that reduced Gram matrix -
    its leading blocks are the factors of every lower degree.  If it is not comfortably positive definite (constant or
    duplicated features, everything clipped, ...) fall back to an eigen-decomposition of each leading block with lstsq's
    rank cut-off."""

    def __init__(self, G: np.ndarray, n: int, F: int, D: int):
        self.G, self.n, self.F, self.D = G, n, F, D
        P = F * (D + 1)
        m = 1 + P - F                                        # reduced column set: column 0 (ones) + the degree >= 1 columns
        R = np.empty((m, m), order="F")                      # four slice copies (np.ix_ gathers took 0.2 ms of the call)
        R[0, 0] = G[0, 0]
        R[0, 1:] = G[0, F:P]
        R[1:, 0] = G[F:P, 0]
        R[1:, 1:] = G[F:P, F:P]
        self.L = None
        self.blocks = {}
        self.eig = {}
        if _lapack is not None:
            c, info = _lapack.dpotrf(R, lower=1, clean=0, overwrite_a=1)
            ok = info == 0
        else:                                                # pragma: no cover
            try:
                c, ok = np.asfortranarray(np.linalg.cholesky(R)), True
            except np.linalg.LinAlgError:
                c, ok = None, False
        if ok:
            dg = c.diagonal()
            if dg.min() > 1e-6 * dg.max():
                self.L = c                                   # lower triangle = the factor (the upper one is not referenced)

    def _block(self, d: int) -> np.ndarray:
        """Leading block of the factor = the factor of degree d's reduced Gram matrix, as its own Fortran-ordered array."""
        if d not in self.blocks:
            m = 1 + self.F * d
            self.blocks[d] = np.asfortranarray(self.L[:m, :m])
        return self.blocks[d]

    def solve(self, d: int, rhs: np.ndarray) -> np.ndarray:
        F = self.F
        Pd = F * (d + 1)
        if self.L is not None:
            Ld = self._block(d)
            b = np.empty(1 + F * d)
            b[0] = rhs[0]
            b[1:] = rhs[F:Pd]
            if _lapack is not None:
                z, _ = _lapack.dtrtrs(Ld, b, lower=1, trans=0)
                z, _ = _lapack.dtrtrs(Ld, z, lower=1, trans=1)
            else:                                            # pragma: no cover
                z = np.linalg.solve(np.tril(Ld).T, np.linalg.solve(np.tril(Ld), b))
            out = np.empty(Pd)
            out[:F] = z[0] / F
            out[F:] = z[1:]
            return out
        if d not in self.eig:
            self.eig[d] = ChebyshevLeastSquares._pinv_factor(self.G[:Pd, :Pd], self.n)
        return ChebyshevLeastSquares._pinv_apply(self.eig[d], rhs)


try:
    from scipy.linalg import lapack as _lapack
except ImportError:                                          # pragma: no cover
    _lapack = None


_BLAS_CTL = None


def _blas_single_thread():
    """The solves between the kernels are small (a few hundred columns): OpenBLAS's thread pool costs more than it gives
    there (and is pathological on an oversubscribed host: 127 ms against 1.3 ms for one evaluate_degree's solves in an
    8-thread container), so they run on the calling thread.  The controller is created once (it scans the loaded libraries)."""
    global _BLAS_CTL
    try:
        if _BLAS_CTL is None:
            from threadpoolctl import ThreadpoolController
            _BLAS_CTL = ThreadpoolController()
        return _BLAS_CTL.limit(limits=1, user_api="blas")
    except Exception:                                        # threadpoolctl missing: keep the library's default
        import contextlib
        return contextlib.nullcontext()


class ChebyshevLeastSquares:
    """The GPU side of evaluate_degree: Gram matrix of [T_0(x) | ... | T_D(x) | y] (qkan_cheb_gram), the small
    minimum-norm solves on the host, explicit residual sums (qkan_cheb_residuals) with one step of iterative
    refinement."""

    def __init__(self, max_degree: int, device: Optional[int] = None, group=None, kernels=None):
        """group: a torch.distributed process group (one process per GPU, NCCL).  Each rank then passes ITS rows to
        solve(); the Gram matrices and the residual sums of the ranks add up, so the only exchange is one all-reduce
        of (P+1)^2 doubles and one of a few hundred - every rank returns the scores of the whole data set.
        kernels: test seam (like ShardedQKANLayer's `compute`): an object with gram(x, y, D) and
        residuals(x, y, w, D, coef, ybar, want_xtr) returning CPU tensors, so that the sharding / reduction / solve logic
        can run under gloo without a GPU; the product path leaves it None and calls libqkan_b200.so."""
        self.group = group
        if not 0 <= max_degree <= 16:
            raise ValueError("the GPU degree evaluation covers 0 <= max_degree <= 16")
        self.kernels = kernels
        self.D = int(max_degree)
        if kernels is not None:
            self.device = torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("qkan_implementation_b200 needs a CUDA device (no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.last = {}

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def gram(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.kernels is not None:
            return self.kernels.gram(x, y, self.D)
        n, F = x.shape
        P = F * (self.D + 1)
        need, slices = ctypes.c_int64(), ctypes.c_int()
        with torch.cuda.device(self.device):
            _b.check(_b.lib().qkan_cheb_gram_workspace(n, F, self.D, ctypes.byref(need), ctypes.byref(slices)))
            ws = torch.empty(max(1, need.value // 8), dtype=torch.float64, device=self.device)
            G = torch.empty((P + 1, P + 1), dtype=torch.float64, device=self.device)
            _b.check(_b.lib().qkan_cheb_gram(x.data_ptr(), y.data_ptr(), n, F, self.D, G.data_ptr(), ws.data_ptr(),
                                             need.value, self._stream()))
        return G

    def residual_sums(self, x, y, w, coef: np.ndarray, ybar: float, want_xtr: bool):
        n, F = x.shape
        D1, P = self.D + 1, F * (self.D + 1)
        if self.kernels is not None:
            sums, tail, xtr = self.kernels.residuals(x, y, w, self.D, coef, ybar, want_xtr)
            return self._reduce_residuals(sums, tail, xtr, want_xtr)
        ctas = ctypes.c_int()
        _b.check(_b.lib().qkan_cheb_residuals_ctas(ctypes.byref(ctas)))
        c = ctas.value
        with torch.cuda.device(self.device):
            cd = torch.from_numpy(np.ascontiguousarray(coef)).to(self.device)
            sums = torch.empty((c, D1, 2), dtype=torch.float64, device=self.device)
            tail = torch.empty((c, 4), dtype=torch.float64, device=self.device)
            xtr = torch.empty((c, D1, P), dtype=torch.float64, device=self.device) if want_xtr else None
            _b.check(_b.lib().qkan_cheb_residuals(x.data_ptr(), y.data_ptr(), w.data_ptr() if w is not None else None, n, F,
                                                  self.D, cd.data_ptr(), float(ybar), sums.data_ptr(), tail.data_ptr(),
                                                  xtr.data_ptr() if want_xtr else None, self._stream()))
            return self._reduce_residuals(sums, tail, xtr, want_xtr)

    def _reduce_residuals(self, sums, tail, xtr, want_xtr):
        """CTA partials are added in CTA order (deterministic), then the ranks' sums (one packed all-reduce)."""
        s = sums.sum(dim=0)
        t = tail.sum(dim=0)
        xr = xtr.sum(dim=0) if want_xtr else None
        if self.group is not None:
            import torch.distributed as dist
            packed = torch.cat([s.reshape(-1), t.reshape(-1)] + ([xr.reshape(-1)] if want_xtr else []))
            dist.all_reduce(packed, group=self.group)
            ns, nt = s.numel(), t.numel()
            s, t = packed[:ns].reshape(s.shape), packed[ns:ns + nt].reshape(t.shape)
            if want_xtr:
                xr = packed[ns + nt:].reshape(xr.shape)
        if self.group is None:                               # one download (one synchronisation) instead of three
            ns, nt = s.numel(), t.numel()
            packed = torch.cat([s.reshape(-1), t.reshape(-1)] + ([xr.reshape(-1)] if want_xtr else [])).cpu().numpy()
            return (packed[:ns].reshape(tuple(s.shape)), packed[ns:ns + nt].reshape(tuple(t.shape)),
                    packed[ns + nt:].reshape(tuple(xr.shape)) if want_xtr else None)
        return s.cpu().numpy(), t.cpu().numpy(), (xr.cpu().numpy() if want_xtr else None)

    @staticmethod
    def _pinv_factor(A: np.ndarray, n: int):
        """Eigen-factor of a leading Gram block for minimum-norm least-squares solves (what np.linalg.lstsq returns
        for the rank-deficient X_d: the T_0 columns of all features are identical).  Eigenvalues below the larger of
        lstsq's own cut-off (rcond = eps max(n, P) on singular values) and the resolution of a Gram matrix are dropped."""
        lam, V = np.linalg.eigh(A)
        eps = np.finfo(np.float64).eps
        P = A.shape[0]
        cut = lam[-1] * max((eps * max(n, P)) ** 2, 64.0 * P * eps)
        keep = lam > cut
        return V[:, keep], lam[keep]

    @staticmethod
    def _pinv_apply(factor, b: np.ndarray) -> np.ndarray:
        Vk, lam = factor
        return Vk @ ((Vk.T @ b) / lam)

    def solve(self, x, y, weights=None, refine: int = 1):
        """x [n, F], y [n], weights [n] or None -> (scores [D+1] = MSE, comp_r2 [D+1]) as evaluate_degree."""
        x = _as_device(x, self.device)
        y = _as_device(y, self.device).reshape(-1)
        w = _as_device(weights, self.device).reshape(-1) if weights is not None else None
        n, F = x.shape
        if y.shape[0] != n or (w is not None and w.shape[0] != n):
            raise ValueError("x, y and weights must have the same number of rows")
        D1, P = self.D + 1, F * (self.D + 1)
        Gd = self.gram(x, y)
        n_all = n
        if self.group is not None:                           # rows are sharded over the ranks: Gram matrices add
            import torch.distributed as dist
            dist.all_reduce(Gd, group=self.group)
            n_all = int(round(float(Gd[0, 0].item())))       # G[0, 0] = sum of T_0^2 = number of rows
        G = Gd.cpu().numpy()
        ybar = G[0, P] / n_all                               # column 0 = T_0 of feature 0 = ones
        n_local, n = n, n_all
        coef = np.zeros((D1, P))
        with _blas_single_thread():
            solver = _NestedSolver(G, n, F, self.D)
            for d in range(D1):
                coef[d, :F * (d + 1)] = solver.solve(d, G[:F * (d + 1), P])
        for _ in range(max(0, refine)):                      # iterative refinement on the explicit residuals
            _, _, xr = self.residual_sums(x, y, w, coef, ybar, True)
            with _blas_single_thread():
                for d in range(D1):
                    coef[d, :F * (d + 1)] += solver.solve(d, xr[d, :F * (d + 1)])
        s, t, _ = self.residual_sums(x, y, w, coef, ybar, False)
        scores, comp_r2 = np.zeros(D1), np.zeros(D1)
        eps = np.finfo(float).eps
        for d in range(D1):
            sse, wsse = s[d]
            if w is not None:                                # DegreeOptimizer.py:291-298 (names as in the reference)
                mse, ss_tot, ss_res = wsse / t[2], wsse, t[1]
            else:                                            # :293-302
                mse, ss_tot, ss_res = sse / n, t[0], sse
            scores[d] = mse
            comp_r2[d] = 0.0 if ss_tot < eps else 1 - ss_tot / ss_res      # :305-309
        self.last = {"coef": coef, "gram": G, "ybar": ybar, "rows": n, "rows_local": n_local}
        return scores, comp_r2

    def features(self, x) -> np.ndarray:
        """[D+1, n, F]: T_d(clip(x)) (DegreeOptimizer._compute_transforms, :96-119)."""
        xd = _as_device(x, self.device)
        n, F = xd.shape
        out = torch.empty((self.D + 1, n, F), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _b.check(_b.lib().qkan_cheb_features(xd.data_ptr(), n, F, self.D, out.data_ptr(), self._stream()))
        return out.cpu().numpy()


class DegreeOptimizer:
    """original_degree_optimizer/DegreeOptimizer.py:13-40 (constructor arguments, defaults and attributes)."""

    def __init__(self, network_shape: List[int], max_degree: int, complexity_weight: float = 0.1,
                 significance_threshold: float = 0.05):
        self.fold_caches = {}                                # BaseOptimizer.py:8-10
        self.network_shape = network_shape
        self.num_layers = len(network_shape) - 1
        self.max_degree = max_degree
        self.complexity_weight = complexity_weight
        self.significance_threshold = significance_threshold
        self.transform_cache = {}
        self.degree_scores = {}
        self.data_same = True
        self.optimal_degrees = None
        self.coefficients = None
        self.feature_means = None
        self.feature_stds = None
        self.qkan_layer: Optional[QKANLayer] = None
        self._lsq: Optional[ChebyshevLeastSquares] = None

    def _engine(self) -> ChebyshevLeastSquares:
        if self._lsq is None or self._lsq.D != self.max_degree:
            self._lsq = ChebyshevLeastSquares(self.max_degree)
        return self._lsq

    # ------------------------------------------------------------------ fit / predict (DegreeOptimizer.py:42-95)
    def fit(self, x_data, y_data, weights=None) -> None:
        self.optimal_degrees = self.optimize_layer(layer_idx=0, x_data=x_data, y_data=y_data, weights=weights)
        feature_data = _as_numpy(x_data)
        self.feature_means = np.mean(feature_data, axis=0)
        self.feature_stds = np.std(feature_data, axis=0) + 1e-8
        self._build_layer()

    def _build_layer(self) -> None:
        """Weights of :63-76: W[d][out_idx * N + in_idx] = 1 where the connection (out_idx, in_idx) has degree d."""
        N, K = self.network_shape[0], self.network_shape[1]
        self.qkan_layer = QKANLayer(N=N, K=K, max_degree=self.max_degree)
        degrees = np.asarray(self.optimal_degrees, dtype=np.int64).reshape(K, N)
        onehot = (degrees.reshape(1, K * N) == np.arange(self.max_degree + 1)[:, None]).astype(np.float64)
        for d in range(self.max_degree + 1):
            self.qkan_layer.mul_step.set_weights(d, onehot[d])

    def predict(self, x_data) -> np.ndarray:
        """:78-95, for the whole batch in one launch (the reference's 2-D call raises in MulStep.py:62-66);
        z-scored inputs beyond [-1, 1] are clipped by the layer, with the reference's warning."""
        if self.qkan_layer is None:
            raise RuntimeError('Not fitted yet')
        feature_data = _as_numpy(x_data)
        normalized_data = (feature_data - self.feature_means) / self.feature_stds
        weights = [self.qkan_layer.mul_step._weights[d] for d in range(self.max_degree + 1)]
        return self.qkan_layer.forward(x=normalized_data, weights=weights, verbose=False)

    # ------------------------------------------------------------------ degree evaluation (:96-181)
    def _compute_transforms(self, feature_data: np.ndarray) -> Dict[int, np.ndarray]:
        t = self._engine().features(feature_data)
        return {d: t[d] for d in range(self.max_degree + 1)}

    def evaluate_degree(self, x_data, y_data, weights=None) -> Tuple[np.ndarray, np.ndarray]:
        """:122-158: (scores = MSE per degree, comp_r2 = the reference's R^2 per degree)."""
        cache_key = str(getattr(x_data, "schema", None))
        if cache_key in self.degree_scores and self.data_same:
            print("Using cached degree scores...")
            return self.degree_scores[cache_key]
        x = x_data.to_numpy() if hasattr(x_data, "to_numpy") else x_data
        scores, comp_r2 = self._engine().solve(x, y_data, weights)
        for d in range(self.max_degree + 1):
            print(f"\nDegree {d}:")
            print(f"MSE: {scores[d]:.8f}")
            print(f"R²:  {comp_r2[d]:.8f}")
        return scores, comp_r2

    def is_degree_definitive(self, scores: np.ndarray) -> Tuple[bool, int]:
        """:159-181: the best (lowest-score) degree is definitive when every other degree is worse by at least
        `significance_threshold`, relative to its own score."""
        sc = np.asarray(scores, dtype=np.float64)
        best = int(np.argmin(sc))
        others = np.delete(sc, best)
        gain = (others - sc[best]) / (others + 1e-10)
        return bool(np.all(gain >= self.significance_threshold)), best

    def optimize_layer(self, layer_idx: int, x_data, y_data, weights, num_reads: int = 1000) -> List[List[int]]:
        """:183-253.  The reference compiles a QUBO and samples it with neal; its objective is a sum over functions of
        sum_d a_d q[i, d] + 10 (sum_d q[i, d] - 1)^2 with the same a_d for every function, so the ground state is
        one-hot at argmin_d a_d for every function - returned here directly (`num_reads` is unused)."""
        input_dim = self.network_shape[layer_idx]
        output_dim = self.network_shape[layer_idx + 1]
        scores, _ = self.evaluate_degree(x_data, y_data, weights)
        is_definitive, definitive_degree = self.is_degree_definitive(scores)
        if is_definitive:                                    # :214-219
            best = definitive_degree
        else:                                                # :221-225
            a = [-(scores[d] - scores[d - 1] if d > 0 else scores[d]) + self.complexity_weight * (d ** 2)
                 for d in range(self.max_degree + 1)]
            best = int(np.argmin(a))
        return [[best for _ in range(input_dim)] for _ in range(output_dim)]

    def optimize_network(self, training_data: Dict[str, np.ndarray], num_reads: int = 1000) -> List[List[List[int]]]:
        """:255-275."""
        return [self.optimize_layer(layer_idx=layer, x_data=training_data[f'layer_{layer}_input'],
                                    y_data=training_data[f'layer_{layer}_output'], weights=None, num_reads=num_reads)
                for layer in range(self.num_layers)]

    def _compute_metrics(self, y_true, y_pred, weights=None) -> Dict[str, float]:
        """:277-312 for explicit predictions (host arithmetic on the caller's arrays, as in the reference).  The
        reference's naming is kept: weighted branch 'ss_tot' = sum w err^2, 'ss_res' = sum w y^2; r2 = 1 - ss_tot / ss_res."""
        yt = np.asarray(y_true).reshape(-1, 1)
        err2 = (yt - np.asarray(y_pred).reshape(-1, 1)) ** 2
        if weights is None:
            mse, ss_tot, ss_res = np.mean(err2), np.sum((yt - np.mean(yt)) ** 2), np.sum(err2)
        else:
            w = np.asarray(weights).reshape(-1, 1)
            mse, ss_tot, ss_res = np.average(err2, weights=w), np.sum(w * err2), np.sum(w * yt ** 2)
        if ss_tot < np.finfo(float).eps:
            print(f"Warning: Total sum of squares ({ss_tot}) near zero - data might be over-normalized")
            return {'mse': float(mse), 'r2': 0.0}
        return {'mse': float(mse), 'r2': float(1 - ss_tot / ss_res)}

    # ------------------------------------------------------------------ state (:313-375)
    _DEFAULT_QUERY = {'n_rows': 100000,
                      'columns': ['date_id', 'responder_6', 'weight'] + [f'feature_{i:02d}' for i in range(79)],
                      'sort_by': 'date_id'}
    _HYPER = ('network_shape', 'max_degree', 'complexity_weight', 'significance_threshold')

    def save_state(self, filename: str, query_params: Dict = None) -> None:
        """Same .npy dictionary as the reference (keys and nesting), so states are interchangeable."""
        state = {k: getattr(self, k) for k in self._HYPER}
        state.update(transform_cache=self.transform_cache, degree_scores=self.degree_scores,
                     query_params=dict(self._DEFAULT_QUERY) if query_params is None else query_params, qkan_params=None)
        if self.qkan_layer is not None:
            state['qkan_params'] = {'weights': [w.copy() for w in self.qkan_layer.mul_step._weights],
                                    'feature_means': self.feature_means.copy(), 'feature_stds': self.feature_stds.copy(),
                                    'optimal_degrees': list(self.optimal_degrees)}
        np.save(filename, state)

    def load_state(self, filename: str, current_query_params: dict) -> None:
        state = np.load(filename, allow_pickle=True).item()
        for k in self._HYPER:
            setattr(self, k, state[k])
        self.num_layers = len(self.network_shape) - 1
        qp = state['qkan_params']
        if qp is not None:
            self.feature_means, self.feature_stds, self.optimal_degrees = qp['feature_means'], qp['feature_stds'], qp['optimal_degrees']
            self.qkan_layer = QKANLayer(N=self.network_shape[0], K=self.network_shape[1], max_degree=self.max_degree)
            for d, w in enumerate(qp['weights']):
                self.qkan_layer.mul_step.set_weights(d, w)
        same = self._validate_query(state['query_params'], current_query_params)
        print("Loading cached computations" if same else "Query changed, clearing caches")
        self.data_same = self.data_same and same
        self.transform_cache = state['transform_cache'] if same else {}
        self.degree_scores = state['degree_scores'] if same else {}

    def _validate_query(self, saved_params: dict, current_query_params: dict) -> bool:
        return all(saved_params[k] == current_query_params[k] for k in ('n_rows', 'columns', 'sort_by'))
