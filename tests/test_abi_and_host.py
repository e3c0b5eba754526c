"""CPU: the C-ABI library loads and exports every symbol include/qkan_b200.h declares; the host
mirror of the reference API validates like the reference; batch sharding logic (gloo, 2 ranks)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "qkan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qkan_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from qkan_implementation_b200 import _binding
    lib = _binding.lib()
    names = _header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qkan_b200.h but not exported"
    assert set(_binding.EXPORTS) == set(names)
    v = [ctypes.c_int() for _ in range(3)]
    lib.qkan_version(*[ctypes.byref(i) for i in v])
    assert (v[0].value, v[1].value) == (0, 1)


def test_abi_argument_errors_without_gpu():
    from qkan_implementation_b200 import _binding as b
    lib = b.lib()
    h = ctypes.c_void_p()
    assert lib.qkan_layer_create(ctypes.byref(h), 0, 4, 3, 0, 0, 1, 0) == b.ERR_BAD_SHAPE
    assert lib.qkan_layer_create(ctypes.byref(h), 4, 4, -1, 0, 0, 1, 0) == b.ERR_BAD_SHAPE
    assert b"positive" in lib.qkan_last_error()
    assert lib.qkan_layer_create(ctypes.byref(h), 4, 4, 64, 0, 0, 0, 0) == b.ERR_UNSUPPORTED   # staged engine: D <= 31
    assert lib.qkan_layer_create(ctypes.byref(h), 4, 4, 4096, 0, 0, 1, 0) == b.ERR_UNSUPPORTED  # block engine: D < 2048
    assert lib.qkan_layer_forward(None, None, 1, None, None, None) == b.ERR_BAD_SHAPE


def test_degree_abi_argument_errors_without_gpu():
    from qkan_implementation_b200 import _binding as b
    lib = b.lib()
    need, slices = ctypes.c_int64(), ctypes.c_int()
    assert lib.qkan_cheb_gram_workspace(774_456, 79, 3, ctypes.byref(need), ctypes.byref(slices)) == 0
    # 79 x 3 + y + the ones column = 239 distinct columns: 4 x 4 tiles, upper triangle; then the 239 x 239 internal matrix
    assert need.value == (slices.value * 10 * 64 * 64 + 239 * 239) * 8 and slices.value >= 1
    assert lib.qkan_cheb_gram_workspace(1000, 5, 0, ctypes.byref(need), ctypes.byref(slices)) == 0
    assert need.value == (slices.value * 1 * 64 * 64 + 6 * 6) * 8                  # degree 0: 5 columns + y
    assert lib.qkan_cheb_gram_workspace(100, 79, 17, ctypes.byref(need), None) == b.ERR_BAD_SHAPE   # D <= 16
    assert lib.qkan_cheb_gram(None, None, 10, 3, 2, None, None, 0, None) == b.ERR_BAD_SHAPE
    assert lib.qkan_cheb_residuals(None, None, None, 10, 3, 2, None, 0.0, None, None, None, None) == b.ERR_BAD_SHAPE
    assert lib.qkan_cheb_features(None, 10, 3, 2, None, None) == b.ERR_BAD_SHAPE
    if not torch.cuda.is_available():
        from qkan_implementation_b200 import DegreeOptimizer
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            DegreeOptimizer([3, 1], 2).evaluate_degree(np.zeros((5, 3)), np.zeros(5))


def _plan_chunks(lib, B, N, K, tile, wave, max_chunks=64):
    cuts = (ctypes.c_int64 * (max_chunks + 1))()
    lib.qkan_plan_host_chunks.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
    n = lib.qkan_plan_host_chunks(B, N, K, tile, wave, cuts, max_chunks)
    return n, list(cuts[:max(n, 0) + 1])


def test_host_chunk_schedule(monkeypatch):
    """Host logic of qkan_layer_forward_host (pure arithmetic, no GPU): the chunk boundaries cover the batch, sit on tile
    multiples, keep about 8 MiB of traffic per chunk with at most 16 chunks, and never cut below one kernel wave."""
    from qkan_implementation_b200 import _binding as b
    lib = b.lib()
    for var in ("QKAN_HOST_CHUNKS", "QKAN_HOST_EDGE_DIV", "QKAN_HOST_CUTS"):
        monkeypatch.delenv(var, raising=False)
    rng = np.random.default_rng(0)
    shapes = [(4, 4, 64, 75_776), (16, 16, 16, 18_944), (784, 10, 16, 18_944), (8, 8, 32, 37_888), (3, 5, 1, 1)]
    batches = [1, 5, 63, 64, 65, 4096, 100_000, 1_000_000, 10_000_019] + [int(v) for v in rng.integers(1, 3_000_000, 20)]
    for (N, K, tile, wave) in shapes:
        for B in batches:
            n, cuts = _plan_chunks(lib, B, N, K, tile, wave)
            assert 1 <= n <= 16, (N, K, B, n)
            assert cuts[0] == 0 and cuts[-1] == B and all(a < c for a, c in zip(cuts, cuts[1:]))
            assert all(c % tile == 0 for c in cuts[1:-1])
            sizes = np.diff(cuts)
            per = max((8 << 20) // ((N + K) * 8), -(-B // 16), wave)       # ~8 MiB of traffic, at most 16 chunks, one wave
            if n > 1:                                          # more than one chunk only when every chunk is at least `per` ...
                assert sizes[:-1].min() >= per and sizes[-1] >= per - n * tile
                assert sizes.max() == sizes[0] and len(set(sizes[:-1])) == 1     # ... and the chunks are even (the last takes the remainder)
                assert sizes[0] < 2 * per + tile
            else:
                assert B < 2 * (per + tile)
    # the BASELINE shapes: 1 M samples of N4 K4 -> 7 chunks of ~9 MB; 100 k samples of N784 K10 -> whole waves, not 64 slivers
    n, cuts = _plan_chunks(lib, 1_000_000, 4, 4, 64, 75_776)
    assert n == 7 and cuts[1] == 142_912
    n, cuts = _plan_chunks(lib, 100_000, 784, 10, 16, 18_944)
    assert n <= 6 and min(np.diff(cuts)[:-1]) >= 18_944
    # tuning aids: a fixed number of uniform chunks, smaller edge chunks, explicit boundaries
    monkeypatch.setenv("QKAN_HOST_CHUNKS", "4")
    n, cuts = _plan_chunks(lib, 1_000_000, 4, 4, 64, 75_776)
    assert n == 4 and cuts[-1] == 1_000_000
    monkeypatch.setenv("QKAN_HOST_CHUNKS", "1000")
    n, cuts = _plan_chunks(lib, 1_000_000, 4, 4, 64, 75_776)
    assert n <= 64 and cuts[-1] == 1_000_000 and all(a < c for a, c in zip(cuts, cuts[1:]))
    monkeypatch.delenv("QKAN_HOST_CHUNKS")
    monkeypatch.setenv("QKAN_HOST_EDGE_DIV", "4")
    n, cuts = _plan_chunks(lib, 1_000_000, 4, 4, 64, 75_776)
    sizes = np.diff(cuts)
    assert sizes[0] == sizes[-1] and sizes[0] * 3 < sizes[1] and cuts[-1] == 1_000_000
    monkeypatch.delenv("QKAN_HOST_EDGE_DIV")
    monkeypatch.setenv("QKAN_HOST_CUTS", "1000,500,70000,2000000")
    n, cuts = _plan_chunks(lib, 100_000, 4, 4, 64, 75_776)
    assert cuts == [0, 1000, 70000, 100_000]
    monkeypatch.delenv("QKAN_HOST_CUTS")
    assert _plan_chunks(lib, 0, 4, 4, 64, 64)[0] == b.ERR_BAD_SHAPE
    assert lib.qkan_plan_host_chunks(10, 4, 4, 64, 64, None, 8) == b.ERR_BAD_SHAPE


def test_no_cpu_fallback_when_no_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qkan_implementation_b200 import QKANLayer, _binding
    layer = QKANLayer(4, 4, 3)
    with pytest.raises(_binding.QkanError):
        layer.forward(np.zeros(4), [np.zeros(16)] * 4)


def test_step_validation_matches_reference():
    from qkan_implementation_b200 import ChebyshevStep, MulStep, QKANLayer
    with pytest.raises(ValueError, match="Degree must be positive integer"):      # ChebyshevStep.py:14
        ChebyshevStep(-1)
    with pytest.raises(ValueError, match="between -1 and 1"):                     # ChebyshevStep.py:26
        ChebyshevStep(1).apply_chebyshev(1.5)
    m = MulStep(2, 4)
    assert m._weights.shape == (3, 4) and m.num_weights == 4
    with pytest.raises(ValueError, match="Degree must be between 0 and 2"):       # MulStep.py:32 / test :236
        m.set_weights(3, np.zeros(4))
    with pytest.raises(ValueError, match="Expected 4 weights, got 3"):            # MulStep.py:34 / test :176
        m.set_weights(0, np.zeros(3))
    with pytest.raises(ValueError, match="Weight magnitudes must be <= 1"):       # MulStep.py:36 / test :170
        m.set_weights(0, np.array([1.5, 0, 0, 0]))
    m.set_weights(1, np.array([1, .5, -.5, -1]))
    assert np.array_equal(m._weights[1], [1, .5, -.5, -1])
    with pytest.raises(ValueError, match="does not match expected size 6 = 3\\*2"):   # MulStep.py:62-66
        m.get_weighted_polynomial_matrix(np.zeros(3), 2, 1)
    layer = QKANLayer(4, 4, 3)
    assert (layer.N, layer.K, layer.max_degree) == (4, 4, 3)
    assert layer.mul_step.num_weights == 16 and layer.cheb_step.degree == 3 and layer.lcu_step.max_degree == 3
    with pytest.raises(ValueError, match="Weight magnitudes"):
        layer.forward(np.zeros(4), [np.full(16, 2.0)] * 4)
    with pytest.raises(ValueError, match="Degree must be between 0 and 3"):
        layer.forward(np.zeros(4), [np.zeros(16)] * 5)
    with pytest.raises(ValueError, match="Expected input dimension 4, got 3"):    # QKANLayer.py:41
        layer.get_intermediate_matrices(np.zeros(3), [np.zeros(16)] * 4)
    with pytest.raises(ValueError, match="Expected 4 weight vectors"):            # QKANLayer.py:43
        layer.get_intermediate_matrices(np.zeros(4), [np.zeros(16)] * 3)
    with pytest.raises(ValueError, match="Expected weight dimension 16"):         # QKANLayer.py:48
        layer.get_intermediate_matrices(np.zeros(4), [np.zeros(15)] * 4)


def test_shard_bounds_cover_batch():
    from qkan_implementation_b200.distributed import shard_bounds, shard_sizes
    for B in (0, 1, 7, 8, 1000, 10_000_001):
        for g in (1, 2, 4, 8):
            edges = [shard_bounds(B, g, r) for r in range(g)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[r][1] == edges[r + 1][0] for r in range(g - 1))
            sz = shard_sizes(B, g)
            assert max(sz) - min(sz) <= 1 and sum(sz) == B


def _gloo_worker(rank, world, port, B, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import qkan_oracle as o
    from qkan_implementation_b200.distributed import ShardedQKANLayer
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    N, K, D = 4, 4, 3
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.uniform(-1, 1, (B, N)))
    W = rng.uniform(-1, 1, (D + 1, N * K))
    compute = lambda xs, w: torch.from_numpy(o.forward_closed_form(xs.numpy(), w, N, K, D)) if xs.shape[0] else torch.zeros((0, K), dtype=torch.float64)
    sh = ShardedQKANLayer(compute=compute)
    full = sh.forward(x, W)
    ref = torch.from_numpy(o.forward_closed_form(x.numpy(), W, N, K, D))
    q.put((rank, bool(torch.equal(full, ref)), tuple(full.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [10, 11, 1])
def test_sharded_forward_gloo_world2(B):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (B, 4) for _, _, shape in res)


class _OracleKernels:
    """CPU stand-in for the two degree-evaluation kernels (same outputs, one 'CTA'), so that the sharding / all-reduce /
    solve logic of ChebyshevLeastSquares runs under gloo without a GPU."""

    @staticmethod
    def _design(x, D):
        from oracle import degree_oracle as do
        tr = do.chebyshev_transforms(x.numpy(), D)
        return np.hstack([tr[d] for d in range(D + 1)])

    def gram(self, x, y, D):
        A = np.hstack([self._design(x, D), y.numpy()[:, None]])
        return torch.from_numpy(A.T @ A)

    def residuals(self, x, y, w, D, coef, ybar, want_xtr):
        X, yv = self._design(x, D), y.numpy()
        wv = w.numpy() if w is not None else np.ones_like(yv)
        F = x.shape[1]
        sums, xtr = np.zeros((1, D + 1, 2)), np.zeros((1, D + 1, X.shape[1]))
        for d in range(D + 1):
            r = yv - X[:, :F * (d + 1)] @ coef[d, :F * (d + 1)]
            sums[0, d] = [np.sum(r * r), np.sum(wv * r * r)]
            xtr[0, d] = X.T @ r
        tail = np.array([[np.sum((yv - ybar) ** 2), np.sum(wv * yv * yv), np.sum(wv), np.sum(yv)]])
        return torch.from_numpy(sums), torch.from_numpy(tail), (torch.from_numpy(xtr) if want_xtr else None)


def _gloo_degree_worker(rank, world, port, weighted, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import degree_oracle as do
    from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares
    from qkan_implementation_b200.distributed import shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, F, D = 1501, 6, 3
    rng = np.random.default_rng(1)
    x = rng.normal(0, 0.6, (n, F))
    y = np.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.05 * rng.normal(size=n)
    w = rng.uniform(0.5, 2, n) if weighted else None
    lo, hi = shard_bounds(n, world, rank)
    eng = ChebyshevLeastSquares(D, group=dist.group.WORLD, kernels=_OracleKernels())
    scores, r2 = eng.solve(x[lo:hi], y[lo:hi], None if w is None else w[lo:hi])
    ref_s, ref_r = do.evaluate_degree(x, y, D, w)                 # the reference algorithm on ALL rows
    ok = bool(np.abs(scores - ref_s).max() <= 1e-12 * ref_s.max() and np.abs(r2 - ref_r).max() <= 1e-9 * max(1, np.abs(ref_r).max()))
    q.put((rank, ok, eng.last["rows"], eng.last["rows_local"], scores.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("weighted", [False, True])
def test_sharded_degree_evaluation_gloo_world2(weighted):
    """Rows sharded over two ranks: Gram matrices and residual sums are all-reduced, every rank gets the scores of the
    whole data set (the kernels are replaced by their oracle, the host logic is the product's)."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_degree_worker, args=(r, 2, port, weighted, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _, _ in res), res
    assert [r[2] for r in res] == [1501, 1501] and sum(r[3] for r in res) == 1501
    assert res[0][4] == res[1][4]                                     # both ranks hold the same scores


@pytest.mark.parametrize("weighted", [False, True])
def test_degree_solve_host_logic_single_process(weighted):
    """ChebyshevLeastSquares.solve without a process group (one packed download per residual pass, nested Cholesky solves,
    one refinement step), kernels replaced by their oracle: scores and R^2 of the reference algorithm."""
    from oracle import degree_oracle as do
    from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares
    n, F, D = 2003, 7, 3
    rng = np.random.default_rng(5)
    x = rng.normal(0, 0.6, (n, F))
    y = np.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.05 * rng.normal(size=n)
    w = rng.uniform(0.5, 2, n) if weighted else None
    eng = ChebyshevLeastSquares(D, kernels=_OracleKernels())
    scores, r2 = eng.solve(x, y, w)
    ref_s, ref_r = do.evaluate_degree(x, y, D, w)
    assert np.abs(scores - ref_s).max() <= 1e-12 * ref_s.max()
    assert np.abs(r2 - ref_r).max() <= 1e-9 * max(1, np.abs(ref_r).max())
    assert eng.last["rows"] == n and eng.last["coef"].shape == (D + 1, F * (D + 1))
