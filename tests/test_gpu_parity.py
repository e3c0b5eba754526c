"""GPU: the CUDA path, called through the C ABI (ctypes) and the drop-in Python layer, against
 (1) golden vectors produced by the unmodified reference, (2) the CPU oracle on seeded inputs,
 (3) size-independent properties at BASELINE.json's full batch sizes.
Tolerances (BASELINE.json north_star): 1e-10 relative for complex128 (and the real-only
representation), 1e-4 relative for complex64 - relative 2-norm per sample, plus an absolute
floor for all-zero outputs (the reference's own zero-input test uses atol 1e-6)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_batches, rel_err
from oracle import qkan_oracle as o

pytestmark = pytest.mark.gpu

RTOL = {"complex128": 1e-10, "real64": 1e-10, "complex64": 1e-4}
ATOL = {"complex128": 1e-13, "real64": 1e-13, "complex64": 2e-6}


def assert_close(got, ref, dtype="complex128"):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    bad = np.abs(got - ref) > ATOL[dtype]
    if bad.any():
        g2, r2 = np.atleast_2d(got), np.atleast_2d(ref)
        rows = np.atleast_2d(bad).any(axis=1)
        assert rel_err(g2[rows], r2[rows]) <= RTOL[dtype], f"max abs err {np.abs(got - ref).max()}"


@pytest.fixture(scope="module")
def Q():
    import qkan_implementation_b200 as q
    assert torch.cuda.is_available()
    return q


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("prep", ["analytic", "gates"])
@pytest.mark.parametrize("dtype", ["complex128", "complex64", "real64"])
@pytest.mark.parametrize("N,K,D,path", golden_batches())
def test_golden_batches(Q, N, K, D, path, dtype, prep):
    g = np.load(path)
    if prep == "gates" and dtype != "complex128" and (N, K, D) not in ((4, 4, 3), (8, 8, 5), (8, 8, 1), (8, 8, 16)):
        pytest.skip("the staged validation engine has reduced-precision kernels only for the whole-register tiles")
    layer = Q.QKANLayer(N, K, D, dtype=dtype, prep=prep)
    out, amps = layer.forward(g["x"], list(g["W"]), return_amplitudes=True)
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert_close(out, g["out"], dtype)
    spec = o.circuit_spec(N, K, D)
    assert_close(amps.real * spec.out_scale, g["out"], dtype)
    assert np.abs(amps.imag).max() == 0.0
    # same call on device tensors, and one sample at a time like the reference
    xd = torch.from_numpy(g["x"]).cuda()
    outd = layer.forward(xd, list(g["W"]))
    assert outd.is_cuda and np.array_equal(outd.cpu().numpy(), out)          # host path == device path, bitwise
    one = layer.forward(g["x"][0], list(g["W"]))
    assert one.shape == (K,) and np.array_equal(one, out[0])


def test_kat_reference_test_vector(Q):
    g = np.load(f"{GOLDEN}/kat_layer_4_4_3.npz")
    layer = Q.QKANLayer(4, 4, 3)
    W = list(g["W"])
    assert_close(layer.forward(g["x"], W), g["out"])
    z = layer.forward(g["x_zero"], W)
    assert np.allclose(z, 0, atol=1e-6) and np.abs(z).max() < 1e-15            # QKANLayer.py:250-252
    assert_close(layer.forward(g["x_boundary"], W), g["out_boundary"])
    assert len(z) == 4 and np.all(np.abs(layer.forward(g["x"], W)) <= 1)       # QKANLayer.py:159-160


def test_intermediate_matrices_and_verbose(Q, capsys):
    g = np.load(f"{GOLDEN}/kat_layer_4_4_3.npz")
    layer = Q.QKANLayer(4, 4, 3)
    m = layer.get_intermediate_matrices(g["x"], list(g["W"]))
    assert m["cheb"][0].shape == (16, 16) and m["weighted"][0].shape == (16, 16) and m["lcu"].shape == (16, 16)
    for d in range(4):
        assert np.abs(np.diag(m["cheb"][d]) - g["cheb_diag"][d]).max() <= 1e-15
        assert np.abs(np.diag(m["weighted"][d]) - g["weighted_diag"][d]).max() <= 1e-15
        assert np.count_nonzero(m["weighted"][d] - np.diag(np.diag(m["weighted"][d]))) == 0
    assert np.abs(np.diag(m["lcu"]) - g["lcu_diag"]).max() <= 1e-15
    assert np.abs(m["reshaped"] - g["reshaped"]).max() <= 1e-15
    assert_close(m["final"], g["final"])
    out = layer.forward(g["x"], list(g["W"]), verbose=True)
    assert_close(out, g["out"])
    txt = capsys.readouterr().out
    assert "QKAN Layer Forward Pass" in txt and "Step 4 (LCU)" in txt
    d = layer.get_intermediate_diagonals(np.stack([g["x"], g["x_boundary"]]))
    assert d["weighted"].shape == (2, 4, 16) and np.abs(d["lcu"][0] - g["lcu_diag"]).max() <= 1e-15


@pytest.mark.parametrize("N,K,D,mode", [(4, 4, 3, "compat"), (8, 8, 1, "compat"), (8, 8, 16, "compat"), (5, 3, 7, "compat"), (16, 16, 8, "compat"),
                                        (4, 4, 20, "compat"), (4, 4, 0, "compat"), (4, 4, 3, "paper"), (6, 5, 9, "paper")])
def test_stage_snapshots_of_the_circuit(Q, N, K, D, mode):
    """QKANLayer.py:52-66 from the SIMULATED circuit: the post-selected block amplitudes after CHEB / SELECT / the degree
    sum (qkan_layer_stage_snapshots) equal the closed form evaluated classically (qkan_layer_diagonals, and NumPy here),
    and summing the lcu snapshot like QKANLayer.py:131-133 reproduces forward()."""
    rng = np.random.default_rng(N * 100 + K * 10 + D)
    B = 19
    x = rng.uniform(-1.1, 1.1, (B, N))
    x[0] = 0.0
    x[1, 0], x[2, 0] = 1.0, -1.0
    W = rng.uniform(-1, 1, (D + 1, N * K))
    layer = Q.QKANLayer(N, K, D, mode=mode)
    layer._set_weights(list(W))
    snap = layer.get_intermediate_diagonals(x, source="circuit")
    ref = layer.get_intermediate_diagonals(x, source="closed_form")
    th = np.arccos(np.clip(x, -1, 1))[:, np.arange(N * K) // K]                 # ChebyshevStep.py:52,64
    cheb_np = np.cos(D * th)
    assert np.abs(snap["cheb"] - cheb_np).max() <= 5e-14 and np.abs(ref["cheb"] - cheb_np).max() <= 5e-14
    for d in range(D + 1):
        c = cheb_np if mode == "compat" else np.cos(d * th)
        assert np.abs(snap["weighted"][:, d] - c * W[d]).max() <= 5e-14        # MulStep.py:72
    assert np.abs(snap["lcu"] - ref["lcu"]).max() <= 5e-14
    out = snap["lcu"].reshape(B, K, N).sum(axis=2) / N                          # QKANLayer.py:131-133
    assert np.abs(out - layer.forward(x, list(W))).max() <= 1e-14
    assert np.abs(snap["final"] - o.forward_closed_form(x, W, N, K, D, mode)).max() <= 1e-13


def test_step_api_known_answers(Q):
    k = np.load(f"{GOLDEN}/kat_steps.npz")
    assert abs(Q.ChebyshevStep(1).apply_chebyshev(0.5) - 0.5) < 1e-15           # ChebyshevStep.py:73
    assert abs(Q.ChebyshevStep(2).apply_chebyshev(0.5) + 0.5) < 1e-15
    assert np.abs(Q.ChebyshevStep(2).transform_diagonal(np.array([0.5, -0.5, 0.0])) - k["t2"]).max() < 1e-15
    assert np.abs(Q.ChebyshevStep(1).create_dilated_chebyshev(np.array([0.5, -0.5]), 2) - k["dil"]).max() < 1e-15
    ms = Q.MulStep(1, 4)
    ms.set_weights(1, np.array([1, .5, -.5, -1]))
    assert np.abs(ms.get_weighted_polynomial_matrix(np.array([.5, -.5]), 2, 1) - k["mul_deg1"]).max() < 1e-15
    ms2 = Q.MulStep(2, 4)
    ms2.set_weights(2, np.array([.5, .5, -.5, -.5]))
    assert np.abs(ms2.get_weighted_polynomial_matrix(np.array([.5, -.5]), 2, 2) - k["mul_deg2"]).max() < 1e-15
    lcu = Q.LCUStep(2).get_combined_matrix(np.array([.5, -.5]), ms2, 2)
    assert np.abs(np.diag(lcu) - np.array([-.25, -.25, .25, .25]) / 3).max() < 1e-15


def test_out_of_range_inputs_clip_and_warn(Q, capsys):
    g = np.load(f"{GOLDEN}/clip_4_4_3.npz")
    layer = Q.QKANLayer(4, 4, 3)
    out = layer.forward(g["x"], list(g["W"]))
    assert_close(out, g["out"])
    assert "Values outside [-1,1] range" in capsys.readouterr().out            # ChebyshevStep.py:48-49
    xd = torch.from_numpy(g["x"]).cuda()
    layer.forward(xd, list(g["W"]))
    n_bad = int((np.abs(g["x"]) > 1 + 1e-8).sum())
    assert layer.out_of_range_count() == n_bad
    assert layer.out_of_range_count() == 0                                      # counter resets
    xn = g["x"].copy()
    xn[0, 0] = np.nan
    assert np.isnan(layer.forward(xn, list(g["W"]))[0]).any()                   # np.clip keeps NaN


# ------------------------------------------------------------------- seeded oracle parity
SHAPES = [(4, 4, 3), (4, 8, 2), (8, 4, 2), (3, 2, 4), (5, 3, 1), (16, 16, 8)] + \
         [(8, 8, D) for D in range(1, 17)] + \
         [(4, 4, 10), (4, 4, 20), (1, 1, 0), (2, 2, 1), (1, 5, 2), (7, 1, 3), (33, 3, 2), (100, 10, 5), (4, 4, 31),
          (6, 5, 7), (12, 3, 9), (3, 11, 13)]          # every degree of the BASELINE configs[4] sweep is its own kernel


@pytest.mark.parametrize("N,K,D", SHAPES)
def test_oracle_parity_seeded(Q, N, K, D):
    rng = np.random.default_rng(17 * N + K + D)
    B = 257                                   # ragged: not a multiple of any CTA tile
    x = rng.uniform(-1, 1, (B, N))
    x[3] = 1.0
    x[4] = -1.0
    x[5] = 0.0
    W = rng.uniform(-1, 1, (D + 1, N * K))
    W[0, 0] = 1.0
    W[-1, -1] = -1.0
    ref = o.forward_closed_form(x, W, N, K, D)
    for dtype in ("complex128", "complex64", "real64"):
        layer = Q.QKANLayer(N, K, D, dtype=dtype)
        assert_close(layer.forward(x, W), ref, dtype)
    layer = Q.QKANLayer(N, K, D, prep="gates")
    assert_close(layer.forward(x, W), ref)
    if N * K <= 64 and D <= 8:
        sv, amps = o.statevector_forward(x[:16], W, N, K, D)
        _, a = Q.QKANLayer(N, K, D).forward(x[:16], W, return_amplitudes=True)
        assert np.abs(a - amps).max() <= 1e-14                                 # post-selected amplitudes


@pytest.mark.parametrize("N,K,D", [(784, 10, 5), (1000, 3, 2), (600, 16, 4), (2000, 1, 3), (513, 7, 16), (900, 33, 2), (100, 10, 5)])
def test_wide_input_rows(Q, N, K, D):
    """Wide input rows run the element-owner kernel (lanes own input elements; an input read by two rows is counted once)."""
    rng = np.random.default_rng(N + 3 * K + D)
    B = 61                                    # ragged
    x = rng.uniform(-1.1, 1.1, (B, N))
    x[2] = 0.0
    x[3, ::2] = 1.0
    W = rng.uniform(-1, 1, (D + 1, N * K))
    ref = o.forward_closed_form(x, W, N, K, D)
    n_bad = int(((x < -1 - 1e-8) | (x > 1 + 1e-8)).sum())
    for dtype in ("complex128", "complex64", "real64"):
        layer = Q.QKANLayer(N, K, D, dtype=dtype)
        y, a = layer.forward(x, W, return_amplitudes=True)
        info = layer.kernel_info()
        assert info["degree_factored"] == 1 and info["element_owner"] == 1 and info["scaled_rotations"] == (2 if D <= 8 else 1), info
        assert_close(y, ref, dtype)
        layer.forward(torch.from_numpy(x).cuda(), W)       # device input: the count stays for the caller
        assert layer.out_of_range_count() == n_bad         # an input shared by two windows is counted once
        spec = o.circuit_spec(N, K, D)
        assert_close(np.real(a) * spec.out_scale, ref, dtype)
    big = Q.QKANLayer(N, K, D)                              # several tiles per CTA, ragged tail
    xb = rng.uniform(-1, 1, (3001, N))
    assert_close(big.forward(xb, W), o.forward_closed_form(xb, W, N, K, D))


@pytest.mark.parametrize("N,K,D", [(4, 4, 3), (8, 8, 5), (3, 2, 4), (5, 3, 1)])
def test_paper_mode(Q, N, K, D):
    rng = np.random.default_rng(N + K + D)
    x = rng.uniform(-1, 1, (65, N))
    W = rng.uniform(-1, 1, (D + 1, N * K))
    layer = Q.QKANLayer(N, K, D, mode="paper")
    assert_close(layer.forward(x, W), o.forward_closed_form(x, W, N, K, D, "paper"))


def test_mnist_shape_single_sample(Q):
    g = np.load(f"{GOLDEN}/batch_784_10_5.npz")
    layer = Q.QKANLayer(784, 10, 5)
    assert_close(layer.forward(g["x"], g["W"]), g["out"])
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (37, 784))
    assert_close(layer.forward(x, g["W"]), o.forward_closed_form(x, g["W"], 784, 10, 5))


def test_empty_and_tiny_batches(Q):
    layer = Q.QKANLayer(4, 4, 3)
    W = np.random.default_rng(0).uniform(-1, 1, (4, 16))
    assert layer.forward(np.zeros((0, 4)), W).shape == (0, 4)
    assert layer.forward(torch.zeros((0, 4), dtype=torch.float64, device="cuda"), W).shape == (0, 4)
    for B in (1, 2, 15, 16, 17, 31, 33):
        x = np.random.default_rng(B).uniform(-1, 1, (B, 4))
        assert_close(layer.forward(x, W), o.forward_closed_form(x, W, 4, 4, 3))


def test_weights_are_stateful_like_reference(Q):
    # forward(weights) overwrites mul_step._weights (QKANLayer.py:124-125 -> MulStep.py:39)
    rng = np.random.default_rng(3)
    layer = Q.QKANLayer(4, 4, 3)
    x = rng.uniform(-1, 1, (8, 4))
    W1, W2 = rng.uniform(-1, 1, (4, 16)), rng.uniform(-1, 1, (4, 16))
    y1 = layer.forward(x, list(W1))
    assert np.array_equal(layer.mul_step._weights, W1)
    y2 = layer.forward(x, list(W2))
    assert not np.array_equal(y1, y2)
    assert_close(y2, o.forward_closed_form(x, W2, 4, 4, 3))
    y3 = layer.forward(x, list(W1[:2]))                 # fewer rows: the rest keep their old values
    Wm = W2.copy()
    Wm[:2] = W1[:2]
    assert_close(y3, o.forward_closed_form(x, Wm, 4, 4, 3))
    # device weights: validated on the GPU
    Wd = torch.from_numpy(W1).cuda()
    yd = layer.forward(torch.from_numpy(x).cuda(), Wd)
    assert np.array_equal(yd.cpu().numpy(), y1)
    with pytest.raises(ValueError, match="Weight magnitudes"):
        layer.forward(torch.from_numpy(x).cuda(), Wd * 3)


def test_degree_optimizer_predict_pattern(Q):
    """The hot path's only in-repo caller: DegreeOptimizer.fit builds one-hot weight vectors
    weights[d][out_idx * N + in_idx] = 1 for the selected degree (DegreeOptimizer.py:63-73) and predict
    passes a 2-D z-scored array (:89-93) - which raises in the reference for more than one row and is
    the batched call here; inputs beyond [-1, 1] are clipped like ChebyshevStep.py:52."""
    rng = np.random.default_rng(21)
    N, K, D, B = 6, 3, 3, 500
    degrees = rng.integers(0, D + 1, size=(K, N))
    weights = [np.zeros(N * K) for _ in range(D + 1)]
    for out_idx in range(K):
        for in_idx in range(N):
            weights[int(degrees[out_idx, in_idx])][out_idx * N + in_idx] = 1.0
    data = rng.normal(0, 1, (B, N))                                   # z-scored features, |x| > 1 occurs
    layer = Q.QKANLayer(N, K, D)
    pred = layer.forward(data, weights, check_range=False)
    assert pred.shape == (B, K)
    assert_close(pred, o.forward_closed_form(data, np.array(weights), N, K, D))
    assert layer.mul_step._weights.shape == (D + 1, N * K)           # what DegreeOptimizer.py:87,327 reads
    again = layer.forward(data, weights, check_range=False)          # unchanged weights: tables are reused
    assert np.array_equal(again, pred)


# ------------------------------------------------------------------ raw C ABI
def test_c_abi_direct(Q):
    b = Q._binding
    lib = b.lib()
    N, K, D, B = 4, 4, 3, 1000
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.uniform(-1, 1, (B, N))).cuda()
    W = torch.from_numpy(rng.uniform(-1, 1, (D + 1, N * K))).cuda()
    out = torch.empty((B, K), dtype=torch.float64, device="cuda")
    rc = lib.qkan_forward(x.data_ptr(), W.data_ptr(), out.data_ptr(), B, N, K, D, 0, 0, None, None)
    assert rc == 0, lib.qkan_last_error()
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), o.forward_closed_form(x.cpu().numpy(), W.cpu().numpy(), N, K, D))
    h = ctypes.c_void_p()
    assert lib.qkan_layer_create(ctypes.byref(h), N, K, D, 0, 0, 1, 0) == 0
    assert lib.qkan_layer_forward(h, x.data_ptr(), B, out.data_ptr(), None, None) == b.ERR_NO_WEIGHTS
    assert lib.qkan_layer_set_weights(h, (W * 2).data_ptr(), 1, 1, None) == b.ERR_WEIGHT_RANGE
    assert lib.qkan_layer_set_weights(h, W.data_ptr(), 1, 1, None) == 0
    assert lib.qkan_layer_forward(h, x.data_ptr(), B, out.data_ptr(), None, None) == 0
    info = b.KernelInfo()
    assert lib.qkan_layer_info(h, ctypes.byref(info)) == 0
    assert info.qubits == 8 and info.flops_survey == 21504 and info.grid > 0
    # CHEB once per evaluated input element, in the sin-weighted basis (D <= 8): s^2 (1 FMA), 2 full passes of
    # 4 MUL + 8 FMA, the pruned pass (4 FMA); SELECT on each of the D + 1 degree copies of every (a, b) (4 FMA)
    assert info.engine == 0 and info.blocks == 64 and info.scaled_rotations == 2 and info.degree_factored == 1
    assert info.direct_rows == 1 and info.cheb_elements == 4
    assert info.flops_exec == 4 * (20 * 2 + 8 + 2) + 64 * 8 and info.fp_inst_exec == 4 * (12 * 2 + 4 + 1) + 64 * 4
    h2 = ctypes.c_void_p()                                   # D > 8: scaled form Ry = gamma M(t)
    assert lib.qkan_layer_create(ctypes.byref(h2), 8, 8, 12, 0, 0, 1, 0) == 0
    assert lib.qkan_layer_info(h2, ctypes.byref(info)) == 0
    assert info.scaled_rotations == 1 and info.flops_exec == 8 * (16 * 11 + 12) + 8 * 8 * 13 * 8 and info.fp_inst_exec == 8 * 8 * 12 + 8 * 8 * 13 * 4
    lib.qkan_layer_destroy(h2)
    assert lib.qkan_layer_info(h, ctypes.byref(info)) == 0
    assert info.flops_per_block_basis == 64 * (16 * 3 + 4)
    lib.qkan_layer_destroy(h)
    assert b.measure_fma_peak(0, True) > 5.0


def test_c_abi_forward_peers_multi_store(Q):
    """fused output gather: the kernel stores each row into every listed buffer at the rank's row offset
    (here two buffers on the same GPU stand in for NVLink peers)"""
    b = Q._binding
    lib = b.lib()
    N, K, D, B = 8, 8, 4, 3001
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.uniform(-1, 1, (B, N))).cuda()
    W = rng.uniform(-1, 1, (D + 1, N * K))
    layer = Q.QKANLayer(N, K, D)
    y = layer.forward(x, W)
    row0, Btot = 517, 5000
    bufs = [torch.full((Btot, K), 7.0, dtype=torch.float64, device="cuda") for _ in range(3)]
    ptrs = (ctypes.c_void_p * 3)(*[ctypes.c_void_p(t.data_ptr()) for t in bufs])
    rc = lib.qkan_layer_forward_peers(layer._engine.handle(), x.data_ptr(), B, ptrs, 3, row0, None)
    assert rc == 0, lib.qkan_last_error()
    torch.cuda.synchronize()
    for t in bufs:
        assert torch.equal(t[row0:row0 + B], y)
        assert float(t[:row0].min()) == 7.0 and float(t[row0 + B:].max()) == 7.0      # nothing else touched
    assert lib.qkan_layer_forward_peers(layer._engine.handle(), x.data_ptr(), B, ptrs, 9, row0, None) == b.ERR_BAD_SHAPE
    gates = Q.QKANLayer(N, K, D, prep="gates")
    gates.forward(x[:4], W)
    assert lib.qkan_layer_forward_peers(gates._engine.handle(), x.data_ptr(), B, ptrs, 1, 0, None) == b.ERR_UNSUPPORTED


# ------------------------------------------------- full-size, size-independent properties
@pytest.mark.parametrize("N,K,D", [(4, 4, 1), (4, 4, 3), (8, 8, 2), (8, 8, 16), (16, 16, 8), (784, 10, 5), (5, 3, 7)])
def test_special_input_values(Q, N, K, D):
    """The scaled-rotation pre-pass (one rsqrt, quarter-turn switch at |x| = 1/sqrt 2) at its edge cases: +-1, +-0,
    denormals, values next to the switch, values next to 1, clipped and infinite inputs."""
    r = np.sqrt(0.5)
    specials = np.array([1.0, -1.0, 0.0, -0.0, 1e-300, -1e-300, 5e-324, r, -r, np.nextafter(r, 1), np.nextafter(r, 0),
                         np.nextafter(1.0, 0), -np.nextafter(1.0, 0), 1.0 + 1e-9, -3.0, np.inf, -np.inf, 0.5, 1e-8, -1e-160])
    rng = np.random.default_rng(N + D)
    B = 64
    x = rng.choice(specials, size=(B, N))
    x[0, :] = specials[np.arange(N) % len(specials)]
    x[1, :] = specials[(np.arange(N) + 7) % len(specials)]
    W = rng.uniform(-1, 1, (D + 1, N * K))
    W.flat[:6] = [1.0, -1.0, 0.0, 1e-300, r, -r][:W.size]
    ref = o.forward_closed_form(x, W, N, K, D)
    assert np.isfinite(ref).all()
    for dtype in ("complex128", "complex64", "real64"):
        y = Q.QKANLayer(N, K, D, dtype=dtype).forward(x, W, check_range=False)
        assert np.isfinite(y).all()
        assert_close(y, ref, dtype)


@pytest.mark.parametrize("env", [{"QKAN_BLOCK_TUNE": "1:256:4:1"}, {"QKAN_BLOCK_TUNE": "1:256:3:2"}, {"QKAN_BLOCK_TUNE": "1:256:2:4"},
                                 {"QKAN_BLOCK_TUNE": "1:128:8:1"}, {"QKAN_BLOCK_TUNE": "4:128:4:1"}, {"QKAN_BLOCK_NO_DT": "1"},
                                 {"QKAN_BLOCK_NO_DIRECT": "1"}, {"QKAN_BLOCK_FORCE_ELEM": "1"}, {"QKAN_BLOCK_NO_ELEM": "1"}, {"QKAN_ELEM_GR": "2"}, {"QKAN_ELEM_GR": "5"},
                                 {"QKAN_BLOCK_FORCE_ELEM": "1", "QKAN_BLOCK_TUNE": "1:256:3:2"}, {"QKAN_BLOCK_FORCE_ELEM": "1", "QKAN_BLOCK_TUNE": "1:256:4:1"},
                                 {"QKAN_BLOCK_STRIDED": "1"}, {"QKAN_BLOCK_STRIDED": "0"}, {"QKAN_BLOCK_SUB": "1"},
                                 {"QKAN_BLOCK_SUB": "2"}, {"QKAN_HOST_PATH": "staged"}, {"QKAN_HOST_PATH": "zero_copy"}, {"QKAN_HOST_PATH": "copy_out"},
                                 {"QKAN_HOST_NO_GRAPH": "1"}])
def test_tuning_variants_agree(Q, env, monkeypatch):
    """Every kernel variant / schedule reachable through the tuning knobs gives the oracle's numbers on ragged batches
    (the knobs are read at layer creation / launch)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for (N, K, D, B) in ((4, 4, 3, 1), (4, 4, 3, 1029), (8, 8, 2, 70_001), (8, 8, 4, 333), (16, 16, 8, 77), (100, 10, 5, 130)):
        rng = np.random.default_rng(B + N)
        x = rng.uniform(-1, 1, (B, N))
        W = rng.uniform(-1, 1, (D + 1, N * K))
        ref = o.forward_closed_form(x, W, N, K, D)
        try:
            layer = Q.QKANLayer(N, K, D)
            y = layer.forward(torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda())
        except Q._binding.QkanError as e:      # a forced (U, CTA, SU) combination is not built for every degree / shape
            assert "QKAN_BLOCK_TUNE" in env and e.code == -2, e
            continue
        assert_close(y.cpu().numpy(), ref)
        xp = torch.from_numpy(x).pin_memory()
        op = torch.empty((B, K), dtype=torch.float64).pin_memory()
        layer.forward(xp.numpy(), W, out=op.numpy(), check_range=False)          # pinned: zero-copy (or staged by env)
        assert_close(op.numpy(), ref)


def test_rejected_device_weights_leave_the_layer_usable(Q):
    """|w| > 1 in a CUDA weight tensor raises like MulStep.set_weights (MulStep.py:36-37) BEFORE anything is overwritten:
    the next forward with the previous (host) weights gives the previous result."""
    rng = np.random.default_rng(3)
    N, K, D = 4, 4, 3
    x = rng.uniform(-1, 1, (33, N))
    W = rng.uniform(-1, 1, (D + 1, N * K))
    layer = Q.QKANLayer(N, K, D)
    ref = layer.forward(x, list(W))
    bad = torch.from_numpy(W * 1.5).cuda()
    with pytest.raises(ValueError, match="Weight magnitudes"):
        layer.forward(torch.from_numpy(x).cuda(), bad)
    assert np.array_equal(layer.mul_step._weights, W)                          # host mirror untouched
    assert np.array_equal(layer.forward(x, list(W)), ref)                      # and the device tables still serve it
    good = torch.from_numpy(W * 0.5).cuda()
    y = layer.forward(torch.from_numpy(x).cuda(), good)
    assert_close(y.cpu().numpy(), o.forward_closed_form(x, W * 0.5, N, K, D))
    assert torch.equal(layer.forward(torch.from_numpy(x).cuda(), good), y)     # same tensor again: upload skipped
    good.mul_(0.5)                                                             # modified in place: uploaded again
    assert_close(layer.forward(torch.from_numpy(x).cuda(), good).cpu().numpy(), o.forward_closed_form(x, W * 0.25, N, K, D))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_device_is_checked_and_current_device_kept(Q):
    """The engine's tables live on one device: tensors of another device are refused, and no call changes the caller's
    current device."""
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, (17, 4))
    W = rng.uniform(-1, 1, (4, 16))
    torch.cuda.set_device(0)
    layer = Q.QKANLayer(4, 4, 3, device=1)
    y = layer.forward(torch.from_numpy(x).to("cuda:1"), list(W))
    assert y.device.index == 1 and torch.cuda.current_device() == 0
    assert_close(y.cpu().numpy(), o.forward_closed_form(x, W, 4, 4, 3))
    assert_close(layer.forward(x, list(W)), o.forward_closed_form(x, W, 4, 4, 3))   # host buffers
    assert torch.cuda.current_device() == 0
    with pytest.raises(ValueError, match="engine lives on cuda:1"):
        layer.forward(torch.from_numpy(x).to("cuda:0"), list(W))


@pytest.mark.parametrize("D", [0, 1, 3, 4, 16, 17])
def test_any_width_creates_and_runs(Q, D):
    """Kernel selection and launch agree on the shared-memory need for every input width (the selection once admitted
    layouts whose launch was rejected around N = 355..383): create + forward over a sweep of N."""
    rng = np.random.default_rng(D)
    for K in (1, 4, 8):
        W0 = rng.uniform(-1, 1, (D + 1, 820 * K))
        x0 = rng.uniform(-1, 1, (5, 820))
        for N in list(range(300, 420, 3)) + list(range(420, 820, 37)):
            W = np.ascontiguousarray(W0[:, :N * K])
            x = np.ascontiguousarray(x0[:, :N])
            y = Q.QKANLayer(N, K, D).forward(x, W)
            assert_close(y, o.forward_closed_form(x, W, N, K, D))


@pytest.mark.parametrize("dtype", ["complex128", "complex64", "real64"])
@pytest.mark.parametrize("N,K,D,B", [(784, 10, 5, 40_037), (300, 7, 3, 50_001), (100, 10, 5, 200_003)])
def test_element_owner_loop_orders_agree(Q, monkeypatch, N, K, D, B, dtype):
    """The element-owner kernel walks wide layers row-outer (all chunks of a CTA through one output row before the next, so the
    row's slice of the SELECT table stays in L1) and everything else chunk-outer: every (sample, row) is evaluated by the same
    code either way, so the two orders must give the same bits - ragged batch, out-of-range inputs and their count included."""
    rng = np.random.default_rng(N + D)
    x = rng.uniform(-1, 1, (B, N))
    x[3, ::5] = 1.0
    x[B - 2, 1::7] = -1.25
    x[B // 2, 0] = 3.0
    W = rng.uniform(-1, 1, (D + 1, N * K))
    n_bad = int(((x < -1 - 1e-8) | (x > 1 + 1e-8)).sum())
    xd, Wd = torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda()
    outs = []
    for order in ("0", "1"):
        monkeypatch.setenv("QKAN_ELEM_ROW_OUTER", order)
        layer = Q.QKANLayer(N, K, D, dtype=dtype)
        y = layer.forward(xd, Wd)
        assert layer.kernel_info()["element_owner"] == 1
        assert layer.out_of_range_count() == n_bad
        outs.append(y)
    assert torch.equal(outs[0], outs[1])
    idx = np.concatenate([np.arange(8), np.arange(B - 8, B), [B // 2]])
    assert_close(outs[1][torch.from_numpy(idx).cuda()].cpu().numpy(), o.forward_closed_form(x[idx], W, N, K, D), dtype)


@pytest.mark.parametrize("N,K,D,B", [(4, 4, 3, 1_000_000),        # BASELINE configs[1]
                                     (16, 16, 8, 1_000_000),     # configs[2]
                                     (784, 10, 5, 100_000),      # configs[3]
                                     (8, 8, 1, 2_000_000), (8, 8, 4, 1_000_000), (8, 8, 16, 1_000_000)])   # configs[4] (per-GPU share)
def test_full_size_properties(Q, N, K, D, B):
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1)
    W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 2 - 1)
    layer = Q.QKANLayer(N, K, D)
    xd, Wd = x.cuda(), W.cuda()
    y = layer.forward(xd, Wd)
    # (1) sampled rows against the oracle
    idx = torch.randint(0, B, (min(4096, max(64, (1 << 22) // (N * K))),), generator=gen)
    assert_close(y[idx.cuda()].cpu().numpy(), o.forward_closed_form(x[idx].numpy(), W.numpy(), N, K, D))
    # (2) determinism and batch-slicing invariance: any slice gives bitwise the same rows
    lo, hi = B // 3 + 1, B // 3 + 1 + 50_001 if B > 100_000 else B // 2
    assert torch.equal(layer.forward(xd[lo:hi].contiguous(), Wd), y[lo:hi])
    assert torch.equal(layer.forward(xd, Wd), y)
    # (3) linearity in the weights: f(x; a W1 + b W2) = a f(x; W1) + b f(x; W2)
    W2 = (torch.rand(W.shape, dtype=torch.float64, generator=gen) * 2 - 1).cuda()
    ya = layer.forward(xd, 0.25 * Wd + 0.5 * W2)
    yb = 0.25 * y + 0.5 * layer.forward(xd, W2)
    assert float((ya - yb).abs().max()) < 1e-13
    # (4) parity of T_D: f(-x) = (-1)^D f(x);  |out| <= 1 (QKANLayer.py:160)
    ym = layer.forward(-xd, Wd)
    assert float((ym - ((-1) ** D) * y).abs().max()) < 1e-13
    assert float(y.abs().max()) <= 1.0
    # (5) checksum against the closed form over the WHOLE batch (float64 on the GPU via torch, chunked)
    src = torch.arange(N * K, device="cuda") // K
    wbar = Wd.mean(dim=0)[None, :]
    step = max(1, (1 << 24) // (N * K))
    worst = 0.0
    for lo5 in range(0, B, step):
        xs = xd[lo5:lo5 + step]
        c = torch.cos(D * torch.arccos(xs.clamp(-1, 1)))[:, src]
        ref = (c * wbar).reshape(xs.shape[0], K, N).sum(dim=2) / N
        worst = max(worst, float((y[lo5:lo5 + step] - ref).abs().max()))
    assert worst < 1e-13
