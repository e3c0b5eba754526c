"""FABLE block-encoding circuits (Camps & Van Beeumen, "Fast Approximate BLock Encodings") as
gate lists, and their evaluation on the GPU.

The reference builds these with the third-party ``fable`` package (un-vendored, unpinned;
call sites ChebyshevStep.py:124, MulStep.py:107, LCUStep.py:60, SUMStep.py:31) and checks
them with Qiskit Aer's ``unitary_simulator`` (``verify_unitary``: MulStep.py:115-166,
LCUStep.py:69-107, SUMStep.py:40-78).  Neither package exists here, so this module restates
the published construction and evaluates the circuits with
``qkan_simulate_circuit`` (libqkan_b200.so).  What the reference's tests pin - and what
``tests/`` check - is the block-encoding identity  ``U[:N, :N] * alpha * 2^n == A``;
the internal gate order of the third-party package is not pinned by anything in the reference.

Qubit labels (little-endian, qubit 0 = least significant bit of the amplitude index), i.e. the
order the circuit has after FABLE's final ``reverse_bits()``:
    system / column register  s_t = qubit t          (t < n)
    row register              r_t = qubit n + t
    flag                            qubit 2n          (most significant)
Circuit:  H on the row register;  O_A = uniformly controlled Ry(2 arccos a_ij) on the flag,
controls (i, j) = (row register, system register), compiled into 4^n Ry + 4^n CX through the
Gray-code / Walsh-Hadamard angle transform;  SWAP(r_t, s_t);  H on the row register.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from . import _binding as _b

H, RY, CX, SWAP, X, Z = 0, 1, 2, 3, 4, 5
_NAMES = {H: "h", RY: "ry", CX: "cx", SWAP: "swap", X: "x", Z: "z"}


def _gray(i: int) -> int:
    return i ^ (i >> 1)


def sfwht(v: np.ndarray) -> np.ndarray:
    """Scaled fast Walsh-Hadamard transform: butterflies with a factor 1/2 per stage."""
    a = np.array(v, dtype=np.float64)
    n = a.shape[0]
    h = 1
    while h < n:
        a = a.reshape(-1, 2, h)
        a = np.stack([(a[:, 0] + a[:, 1]) * 0.5, (a[:, 0] - a[:, 1]) * 0.5], axis=1).reshape(n)
        h *= 2
    return a


def gray_permute(v: np.ndarray) -> np.ndarray:
    idx = np.arange(v.shape[0])
    return v[idx ^ (idx >> 1)]


@dataclass
class FableCircuit:
    """A gate list on ``num_qubits = 2n + 1`` qubits; stands in for the qiskit.QuantumCircuit the
    reference's ``fable(A, 0)`` returns."""
    n: int
    gates: List[Tuple[int, int, int]] = field(default_factory=list)     # (kind, q0, q1)
    params: List[float] = field(default_factory=list)

    @property
    def num_qubits(self) -> int:
        return 2 * self.n + 1

    def add(self, kind: int, q0: int, q1: int = 0, theta: float = 0.0):
        self.gates.append((kind, q0, q1))
        self.params.append(theta)

    def count_ops(self) -> dict:
        out = {}
        for k, _, _ in self.gates:
            out[_NAMES[k]] = out.get(_NAMES[k], 0) + 1
        return out

    def size(self) -> int:
        return len(self.gates)

    def depth(self) -> int:
        """Circuit depth as qiskit.QuantumCircuit.depth() counts it: the longest chain of gates that share a qubit."""
        level = [0] * self.num_qubits
        for kind, q0, q1 in self.gates:
            qs = (q0, q1) if kind in (CX, SWAP) else (q0,)
            d = 1 + max(level[q] for q in qs)
            for q in qs:
                level[q] = d
        return max(level) if level else 0

    # -------------------------------------------------------------- GPU evaluation
    def columns(self, basis_states) -> np.ndarray:
        """Evolve |j> for every j in ``basis_states`` on the GPU; returns complex128 [len, 2^num_qubits]."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("qkan_implementation_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
        basis = torch.as_tensor(np.asarray(basis_states, dtype=np.int64), device=dev)
        g = torch.as_tensor(np.asarray(self.gates, dtype=np.int32).reshape(-1, 3), device=dev).contiguous()
        p = torch.as_tensor(np.asarray(self.params, dtype=np.float64), device=dev)
        if g.numel() == 0:                       # keep the pointers valid for an empty circuit
            g = torch.zeros((1, 3), dtype=torch.int32, device=dev)
            p = torch.zeros((1,), dtype=torch.float64, device=dev)
        out = torch.empty((basis.shape[0], 1 << self.num_qubits), dtype=torch.complex128, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _b.check(_b.lib().qkan_simulate_circuit(g.data_ptr(), p.data_ptr(), len(self.gates), self.num_qubits,
                                                basis.data_ptr(), basis.shape[0], out.data_ptr(), stream))
        return out.cpu().numpy()

    def block(self, size: int | None = None) -> np.ndarray:
        """Top-left block of the circuit's unitary (all ancillas |0> in, |0> out): U[:size, :size]."""
        N = 1 << self.n
        size = N if size is None else size
        cols = self.columns(np.arange(size))                # row j = U |j>
        return cols[:, :size].T                             # U[i, j] = <i| U |j>


def fable(A: np.ndarray, eps: float = 0.0):
    """Block-encode the real matrix A: returns ``(circuit, alpha)`` with
    ``circuit.block() * alpha * 2^n == A`` (zero padded to 2^n).  ``eps`` = FABLE's angle
    threshold (the reference always passes 0: exact)."""
    A = np.array(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("fable expects a square real matrix")
    dim = A.shape[0]
    n = max(1, int(np.ceil(np.log2(dim)))) if dim > 1 else 1
    N = 1 << n
    alpha = float(np.max(np.abs(A))) if A.size else 0.0
    if alpha > 1.0:
        alpha = alpha + np.sqrt(np.finfo(np.float64).eps)
        A = A / alpha
    else:
        alpha = 1.0
    Ap = np.zeros((N, N))
    Ap[:dim, :dim] = A
    theta = gray_permute(sfwht(2.0 * np.arccos(Ap.reshape(-1))))      # control index = i * N + j (row major)

    circ = FableCircuit(n)
    flag = 2 * n
    # control index bit p: p < n -> system qubit p (j); p >= n -> row qubit n + (p - n) (i): qubit p itself
    for t in range(n):
        circ.add(H, n + t)
    nctl = 2 * n
    parity = 0                                     # CX gates pending since the last emitted rotation
    for i in range(1 << nctl):
        if abs(theta[i]) > eps:
            for pbit in range(nctl):               # flush pending CXs (equal pairs cancel)
                if (parity >> pbit) & 1:
                    circ.add(CX, pbit, flag)
            parity = 0
            circ.add(RY, flag, 0, float(theta[i]))
        if i + 1 < (1 << nctl):
            ctl = (_gray(i) ^ _gray(i + 1)).bit_length() - 1
        else:
            ctl = nctl - 1                         # closes the Gray cycle
        parity ^= 1 << ctl
    for pbit in range(nctl):
        if (parity >> pbit) & 1:
            circ.add(CX, pbit, flag)
    for t in range(n):
        circ.add(SWAP, n + t, t)
    for t in range(n):
        circ.add(H, n + t)
    return circ, alpha


def verify_unitary(circuit: FableCircuit, expected_matrix: np.ndarray, scale: float) -> float:
    """What the reference's tests compute (MulStep.py:115-166): relative Frobenius error of
    ``U[:n, :n] * scale * n`` against ``expected_matrix`` (n = matrix size)."""
    n = expected_matrix.shape[0]
    top_left = circuit.block()[:n, :n].real * scale * (1 << circuit.n)
    den = np.linalg.norm(expected_matrix)
    diff = np.linalg.norm(top_left - expected_matrix)
    return float(diff / den) if den > 1e-10 else float(diff)
