"""CPU: the oracle against the reference's golden vectors and known answers."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_batches, rel_err
from oracle import qkan_oracle as o


@pytest.mark.parametrize("N,K,D,path", golden_batches())
def test_closed_form_matches_reference(N, K, D, path):
    g = np.load(path)
    out = o.forward_closed_form(g["x"], g["W"], N, K, D)
    assert np.abs(out - g["out"]).max() <= 4e-16


@pytest.mark.parametrize("N,K,D,path", [t for t in golden_batches() if t[0] * t[1] <= 256])
def test_reference_style_and_statevector_match_reference(N, K, D, path):
    g = np.load(path)
    rs = np.stack([o.forward_reference_style(x, list(g["W"]), N, K, D) for x in g["x"][:4]])
    assert np.abs(rs - g["out"][:4]).max() <= 4e-16
    sv, amps = o.statevector_forward(g["x"], g["W"], N, K, D)
    assert np.abs(sv - g["out"]).max() <= 2e-15
    assert np.abs(amps.imag).max() == 0.0          # every gate is real
    sv32, _ = o.statevector_forward(g["x"], g["W"], N, K, D, dtype=np.complex64)
    assert np.abs(sv32 - g["out"]).max() <= 1e-5


def test_kat_layer():
    g = np.load(f"{GOLDEN}/kat_layer_4_4_3.npz")
    # Appendix D, KAT-1..3 (values printed in SURVEY.md / BASELINE.md)
    assert np.allclose(g["out"], [-0.040779429890372136, -0.0431279018067088, 0.1953124009009744,
                                  -0.040911180772127104], rtol=0, atol=1e-16)
    W = g["W"]
    for xk, ok in (("x", "out"), ("x_zero", "out_zero"), ("x_boundary", "out_boundary")):
        assert np.abs(o.forward_closed_form(g[xk], W, 4, 4, 3)[0] - g[ok]).max() <= 2e-16
        assert np.abs(o.statevector_forward(g[xk], W, 4, 4, 3)[0][0] - g[ok]).max() <= 1e-15
    assert np.allclose(g["out_zero"], 0, atol=1e-6)                   # QKANLayer.py:250-252
    assert np.array_equal(g["out_verbose"], g["out"])
    inter = o.intermediate_matrices(g["x"], list(W), 4, 4, 3)
    for d in range(4):
        assert np.abs(np.diag(inter["cheb"][d]) - g["cheb_diag"][d]).max() == 0
        assert np.abs(np.diag(inter["weighted"][d]) - g["weighted_diag"][d]).max() == 0
    assert np.abs(np.diag(inter["lcu"]) - g["lcu_diag"]).max() == 0
    assert np.abs(inter["reshaped"] - g["reshaped"]).max() == 0
    assert np.abs(inter["final"] - g["final"]).max() == 0
    sd = o.stage_diagonals(g["x"], W, 4, 4, 3)
    assert np.abs(sd["lcu"][0] - g["lcu_diag"]).max() <= 2e-16
    assert np.abs(sd["reshaped"][0] - g["reshaped"]).max() <= 2e-16


def test_kat_steps():
    g = np.load(f"{GOLDEN}/kat_steps.npz")
    assert np.allclose(o.chebyshev_values(np.array([0.5]), 1), 0.5)                      # ChebyshevStep.py:73-76
    assert np.allclose(o.chebyshev_values(np.array([0.5, -0.5, 0.0]), 2), g["t2"])       # :83-91
    assert np.allclose(g["t2"], [-0.5, -0.5, -1.0])
    assert np.allclose(o.dilated_chebyshev_matrix(np.array([0.5, -0.5]), 2, 1), g["dil"])  # :93-102
    assert np.allclose(np.diag(g["mul_deg1"]), [.5, .25, .25, .5])                       # MulStep.py:192-200
    assert np.allclose(np.diag(g["mul_deg2"]), [-.25, -.25, .25, .25])                   # MulStep.py:217-225
    # SUM known answer (SUMStep.py:86-94): diag [1,.5,-.5,-1], N=K=2 -> [.75,-.75]
    d = np.array([1, .5, -.5, -1.0])
    assert np.allclose(np.sum(d.reshape(2, 2, order="F"), axis=0) / 2, [0.75, -0.75])


def test_clip_and_range_warning():
    g = np.load(f"{GOLDEN}/clip_4_4_3.npz")
    assert np.abs(o.forward_closed_form(g["x"], g["W"], 4, 4, 3) - g["out"]).max() <= 4e-16
    assert np.abs(o.statevector_forward(g["x"], g["W"], 4, 4, 3)[0] - g["out"]).max() <= 2e-15
    assert len(o.out_of_range(g["x"])) > 0 and int(g["n_warnings"]) > 0


def test_weight_validation_messages():
    with pytest.raises(ValueError, match="Weight magnitudes must be <= 1"):
        o.forward_reference_style(np.zeros(2), [np.array([1.5, 0, 0, 0])] * 2, 2, 2, 1)
    with pytest.raises(ValueError, match="Expected 4 weights, got 3"):
        o.forward_reference_style(np.zeros(2), [np.zeros(3)] * 2, 2, 2, 1)
    with pytest.raises(ValueError, match="does not match"):
        o.forward_reference_style(np.zeros(3), [np.zeros(4)] * 2, 2, 2, 1)


def test_paper_mode_differs_and_circuit_agrees():
    g = np.load(f"{GOLDEN}/kat_layer_4_4_3.npz")
    p = o.forward_closed_form(g["x"], g["W"], 4, 4, 3, mode="paper")[0]
    assert np.allclose(p, [0.09147771, -0.03983563, -0.07082646, -0.07507989], atol=1e-8)   # Appendix D
    sv, _ = o.statevector_forward(g["x"], g["W"], 4, 4, 3, mode="paper")
    assert np.abs(sv[0] - p).max() <= 1e-15


def test_spec_flop_counts():
    # SURVEY.md 8(d) table
    assert o.circuit_spec(4, 4, 3).flops_complex == 21504
    assert o.circuit_spec(16, 16, 8).flops_complex == 2850816
    assert o.circuit_spec(784, 10, 5).flops_complex == 113246208
    assert o.circuit_spec(8, 8, 1).flops_complex == 39936
    assert o.circuit_spec(8, 8, 16).flops_complex == 1769472
