"""CPU oracle for the QKANLayer.forward hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``qkan_implementation_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  Every function here is checked (tests/test_oracle.py)
against ``tests/golden/*.npz``, which ``oracle/gen_golden.py`` produced by
importing the UNMODIFIED reference from ``/root/reference`` (fable / qiskit /
qiskit_aer replaced by empty import stubs, which ``forward`` never touches) and
against the known answers the reference's own unit tests hold
(SURVEY.md Appendix D).  What stays unpinned: the north-star's "Qiskit
Statevector" - neither Qiskit nor fable is installable here and the reference
never builds a circuit for ``forward``; the gate list below (`circuit_spec`) is
defined by this project and is pinned only through its post-selected amplitudes
equalling the reference's ``forward`` value.

Three restatements of the same path, each citing the reference lines it follows
(paths relative to /root/reference/QKAN_Steps_original/):

* ``forward_reference_style``  - one sample, dense ``np.diag`` algebra in the
  same order of operations as the reference (used as the timed CPU "port").
* ``forward_closed_form``      - batched closed form.
* ``statevector_forward``      - gate-level complex128 simulation of the
  DILATE/CHEB/MUL/LCU/SUM circuit, the thing the CUDA
  kernel executes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

EPS_RANGE = 1e-8  # ChebyshevStep.py:26,40  (tolerance of the range warning)


# --------------------------------------------------------------------------
# 1. reference-style restatement (one sample, dense matrices)
# --------------------------------------------------------------------------
def chebyshev_values(x: np.ndarray, degree: int) -> np.ndarray:
    """T_degree of every entry after clipping to [-1, 1].

    Follows ChebyshevStep.py:32-53 (transform_diagonal: clip at :52, then the
    per-element apply_chebyshev :18-30 = cos(degree * arccos(x))).  The range
    warning print (:47-49) is not part of the arithmetic and is left out.
    """
    xc = np.clip(np.asarray(x, dtype=np.float64), -1.0, 1.0)
    return np.cos(degree * np.arccos(xc))


def chebyshev_values_elementwise(x: np.ndarray, degree: int) -> np.ndarray:
    """Same values, evaluated one element at a time with the per-element range check and clip
    of apply_chebyshev (ChebyshevStep.py:18-30, called from the list comprehension at :53).
    Used by the timed CPU port so that it has the reference's cost profile."""
    xc = np.clip(np.asarray(x, dtype=np.float64), -1, 1)
    vals = []
    for xi in xc:
        if not np.all((xi >= -1 - EPS_RANGE) & (xi <= 1 + EPS_RANGE)):
            raise ValueError("Input value must be between -1 and 1.")
        vals.append(np.cos(degree * np.arccos(np.clip(xi, -1, 1))))
    return np.array(vals)


def out_of_range(x: np.ndarray) -> np.ndarray:
    """Entries the reference would list in its warning (ChebyshevStep.py:46-49)."""
    x = np.asarray(x, dtype=np.float64)
    return x[~(-1 - EPS_RANGE <= x) | ~(x <= 1 + EPS_RANGE)]


def dilated_chebyshev_matrix(x: np.ndarray, K: int, degree: int, elementwise: bool = False) -> np.ndarray:
    """diag(repeat(T_degree(x), K))  -  ChebyshevStep.py:55-65."""
    vals = chebyshev_values_elementwise(x, degree) if elementwise else chebyshev_values(x, degree)
    return np.diag(np.repeat(vals, K))


def validate_weights(W: np.ndarray, D: int, NK: int) -> None:
    """The checks of MulStep.set_weights (MulStep.py:32-37), same messages."""
    if len(W) > D + 1:
        raise ValueError(f"Degree must be between 0 and {D}")
    for w in W:
        if len(w) != NK:
            raise ValueError(f"Expected {NK} weights, got {len(w)}")
        if not np.all(np.abs(w) <= 1):
            raise ValueError("Weight magnitudes must be <= 1 for unitarity")


def forward_reference_style(x: np.ndarray, weights, N: int, K: int, D: int,
                            mode: str = "compat") -> np.ndarray:
    """One sample through dense diagonal matrices, the way the reference does it.

    QKANLayer.py:122-135 (fast path) -> LCUStep.py:32-36 (zeros + sequential
    accumulation of matrix/(D+1)) -> MulStep.py:59-72 (dilated Chebyshev matrix
    rebuilt for every term, always with degree D - the "degree quirk",
    MulStep.py:20,59 - times the weight row) -> reshape(N, K, 'F'), sum, /N.
    ``mode='paper'`` uses T_d for term d instead (not what the reference does).
    """
    x = np.asarray(x, dtype=np.float64)
    W = np.zeros((D + 1, N * K))
    validate_weights(weights, D, N * K)
    for d, w in enumerate(weights):
        W[d] = w
    if N * K != len(x) * K:  # MulStep.py:62-66
        raise ValueError(f"Weight vector size {N * K} does not match "
                         f"expected size {len(x) * K} = {len(x)}*{K}")
    acc = np.zeros((N * K, N * K))
    for d in range(D + 1):
        cheb = dilated_chebyshev_matrix(x, K, D if mode == "compat" else d, elementwise=True)
        acc += np.diag(np.diag(cheb) * W[d]) / (D + 1)
    return np.sum(np.diag(acc).reshape(N, K, order="F"), axis=0) / N


def intermediate_matrices(x, weights, N, K, D):
    """Dense by-products of QKANLayer.get_intermediate_matrices (QKANLayer.py:30-75)."""
    if len(x) != N:
        raise ValueError(f"Expected input dimension {N}, got {len(x)}")
    if len(weights) != D + 1:
        raise ValueError(f"Expected {D + 1} weight vectors")
    for w in weights:
        if len(w) != N * K:
            raise ValueError(f"Expected weight dimension {N * K}")
    validate_weights(weights, D, N * K)
    W = np.asarray(weights, dtype=np.float64)
    res = {"input": np.asarray(x)}
    res["cheb"] = {d: dilated_chebyshev_matrix(x, K, D) for d in range(D + 1)}
    res["weighted"] = {d: np.diag(np.diag(res["cheb"][d]) * W[d]) for d in range(D + 1)}
    lcu = np.zeros((N * K, N * K))
    for d in range(D + 1):
        lcu += res["weighted"][d] / (D + 1)
    res["lcu"] = lcu
    res["reshaped"] = np.diag(lcu).reshape(N, K, order="F")
    res["final"] = np.sum(res["reshaped"], axis=0) / N
    return res


# --------------------------------------------------------------------------
# 2. closed form, batched  
# --------------------------------------------------------------------------
def forward_closed_form(x: np.ndarray, W: np.ndarray, N: int, K: int, D: int,
                        mode: str = "compat") -> np.ndarray:
    """out[s, b] = (1/N) sum_a lcu[s, a + N b],  lcu[i] = sum_d c_d[i // K] W[d, i] / (D+1).

    Index conventions: dilation is input-major (ChebyshevStep.py:64, value index
    i // K) while the SUM reshape is column-major (QKANLayer.py:132, i = a + N b).
    Accumulates over d in degree order like LCUStep.py:34-36.
    """
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    W = np.asarray(W, dtype=np.float64).reshape(D + 1, N * K)
    B = x.shape[0]
    xc = np.clip(x, -1.0, 1.0)
    th = np.arccos(xc)
    src = np.arange(N * K) // K
    lcu = np.zeros((B, N * K))
    for d in range(D + 1):
        c = np.cos((D if mode == "compat" else d) * th)
        lcu += c[:, src] * W[d][None, :] / (D + 1)
    return lcu.reshape(B, K, N).sum(axis=2) / N


# --------------------------------------------------------------------------
# 3. the circuit and its statevector simulation  
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class CircuitSpec:
    """Register layout, little-endian: deg[l] | f_x | f_w | a[n_a] | b[n_b]."""
    N: int
    K: int
    D: int
    n_a: int
    n_b: int
    l: int

    @property
    def m(self):
        return self.n_a + self.n_b

    @property
    def qubits(self):
        return self.l + 2 + self.m

    @property
    def S(self):
        return 1 << self.qubits

    @property
    def passes(self):
        """P of SURVEY.md 8(d): (D+1) multiplexor passes + Hadamard passes."""
        return (self.D + 1) + (self.m + 2 * self.l + self.n_a)

    @property
    def flops_complex(self):
        return 6 * self.S * self.passes

    def bit_deg(self, i):
        return i

    @property
    def bit_fx(self):
        return self.l

    @property
    def bit_fw(self):
        return self.l + 1

    def bit_a(self, i):
        return self.l + 2 + i

    def bit_b(self, i):
        return self.l + 2 + self.n_a + i

    @property
    def out_scale(self):
        """amp[b] * out_scale = forward value (Appendix C step 5)."""
        Np, Kp, Lp = 1 << self.n_a, 1 << self.n_b, 1 << self.l
        return Np * math.sqrt(Kp) * Lp / (self.N * (self.D + 1))


def _clog2(n: int) -> int:
    return 0 if n <= 1 else (n - 1).bit_length()


def circuit_spec(N: int, K: int, D: int) -> CircuitSpec:
    return CircuitSpec(N, K, D, _clog2(N), _clog2(K), _clog2(D + 1))


def angle_tables(spec: CircuitSpec, x: np.ndarray, W: np.ndarray):
    """cos(theta/2) for the two multiplexors, padded entries = 0 (theta = pi).

    cx[s, b, a] = clip(x[s, (a + N b) // K]);  cw[d, b, a] = W[d, a + N b].
    sin(theta/2) = sqrt(1 - cos^2) >= 0 because theta/2 = arccos(.) in [0, pi].
    """
    N, K, D = spec.N, spec.K, spec.D
    Np, Kp, Lp = 1 << spec.n_a, 1 << spec.n_b, 1 << spec.l
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    W = np.asarray(W, dtype=np.float64).reshape(D + 1, N * K)
    a = np.arange(N)[None, :]
    b = np.arange(K)[:, None]
    i = a + N * b
    cx = np.zeros((x.shape[0], Kp, Np))
    cx[:, :K, :N] = np.clip(x, -1, 1)[:, i // K]
    cw = np.zeros((Lp, Kp, Np))
    cw[: D + 1, :K, :N] = W[:, i]
    return cx, cw


def _hadamard(psi: np.ndarray, bit: int) -> np.ndarray:
    B, S = psi.shape
    v = psi.reshape(B, S >> (bit + 1), 2, 1 << bit)
    out = np.empty_like(v)
    out[:, :, 0] = (v[:, :, 0] + v[:, :, 1]) * (1 / math.sqrt(2))
    out[:, :, 1] = (v[:, :, 0] - v[:, :, 1]) * (1 / math.sqrt(2))
    return out.reshape(B, S)


def _mux_ry(psi: np.ndarray, bit: int, c: np.ndarray, s: np.ndarray) -> np.ndarray:
    """Block-diagonal pass: Ry with per-amplitude-pair (c, s) = cos/sin(theta/2).

    c, s are [B or 1, S] arrays indexed by the full amplitude index (they only
    depend on the control bits, so the two members of a pair see the same value).
    """
    B, S = psi.shape
    v = psi.reshape(B, S >> (bit + 1), 2, 1 << bit)
    cc = np.broadcast_to(c, (B, S)).reshape(B, S >> (bit + 1), 2, 1 << bit)[:, :, 0]
    ss = np.broadcast_to(s, (B, S)).reshape(B, S >> (bit + 1), 2, 1 << bit)[:, :, 0]
    out = np.empty_like(v)
    out[:, :, 0] = cc * v[:, :, 0] - ss * v[:, :, 1]
    out[:, :, 1] = ss * v[:, :, 0] + cc * v[:, :, 1]
    return out.reshape(B, S)


def statevector_forward(x: np.ndarray, W: np.ndarray, N: int, K: int, D: int,
                        mode: str = "compat", dtype=np.complex128,
                        return_state: bool = False):
    """Simulate the Appendix-C gate list for every row of x.

    Returns (out[B, K] float, amps[B, K] complex) - ``amps`` are the amplitudes
    post-selected on deg = f_x = f_w = a = 0, ``out = Re(amps) * spec.out_scale``.
    The Z . UCRy(-theta) . Z of the odd CHEB applications is applied literally
    (two sign passes around a rotation by -theta).
    """
    spec = circuit_spec(N, K, D)
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    B, S = x.shape[0], spec.S
    cx, cw = angle_tables(spec, x, W)
    idx = np.arange(S)
    deg = idx & ((1 << spec.l) - 1)
    a = (idx >> spec.bit_a(0)) & ((1 << spec.n_a) - 1)
    b = (idx >> spec.bit_b(0)) & ((1 << spec.n_b) - 1)
    rdt = np.float64 if dtype == np.complex128 else np.float32
    CX = cx[:, b, a].astype(rdt)                       # [B, S]
    SX = np.sqrt((1 - CX) * (1 + CX)).astype(rdt)
    CW = cw[deg, b, a][None, :].astype(rdt)            # [1, S]
    SW = np.sqrt((1 - CW) * (1 + CW)).astype(rdt)

    psi = np.zeros((B, S), dtype=dtype)
    psi[:, 0] = 1
    # 1. SUM pre-layer / read-out superposition on a, b;  PREPARE on deg
    for i in range(spec.n_a):
        psi = _hadamard(psi, spec.bit_a(i))
    for i in range(spec.n_b):
        psi = _hadamard(psi, spec.bit_b(i))
    for i in range(spec.l):
        psi = _hadamard(psi, spec.bit_deg(i))
    # 2. CHEB: D applications of the input block-encoding, alternating U, Z U^dagger Z
    zsign = np.where((idx >> spec.bit_fx) & 1, -1.0, 1.0).astype(rdt)
    for j in range(D):
        if mode == "paper":
            on = (deg >= j + 1)
            c = np.where(on[None, :], CX, 1).astype(rdt)
            s = np.where(on[None, :], SX, 0).astype(rdt)
        else:
            c, s = CX, SX
        if j % 2 == 0:
            psi = _mux_ry(psi, spec.bit_fx, c, s)
        else:
            zs = np.where(on, zsign, 1).astype(rdt) if mode == "paper" else zsign
            psi = psi * zs
            psi = _mux_ry(psi, spec.bit_fx, c, -s)
            psi = psi * zs
    # 3. MUL / SELECT
    psi = _mux_ry(psi, spec.bit_fw, CW, SW)
    # 4. UNPREPARE, SUM
    for i in range(spec.l):
        psi = _hadamard(psi, spec.bit_deg(i))
    for i in range(spec.n_a):
        psi = _hadamard(psi, spec.bit_a(i))
    # 5. post-selection
    amps = psi[:, (np.arange(K) << spec.bit_b(0))]
    out = amps.real.astype(np.float64) * spec.out_scale
    if return_state:
        return out, amps, psi
    return out, amps


def stage_diagonals(x: np.ndarray, W: np.ndarray, N: int, K: int, D: int):
    """Per-stage diagonals of get_intermediate_matrices for a batch (closed form).

    Returns dict with cheb[B, NK], weighted[B, D+1, NK], lcu[B, NK],
    reshaped[B, N, K], final[B, K]  (QKANLayer.py:52-73, diagonals only).
    """
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    W = np.asarray(W, dtype=np.float64).reshape(D + 1, N * K)
    src = np.arange(N * K) // K
    cheb = chebyshev_values(x, D)[:, src]
    weighted = cheb[:, None, :] * W[None]
    lcu = np.zeros_like(cheb)
    for d in range(D + 1):
        lcu += weighted[:, d] / (D + 1)
    reshaped = lcu.reshape(-1, K, N).transpose(0, 2, 1)
    return {"cheb": cheb, "weighted": weighted, "lcu": lcu,
            "reshaped": reshaped, "final": reshaped.sum(axis=1) / N}
