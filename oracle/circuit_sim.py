"""Dense NumPy gate-list simulator.  TEST INFRASTRUCTURE ONLY (see oracle/qkan_oracle.py header).

Restates what the reference's unit tests obtain from Qiskit Aer's unitary_simulator
(MulStep.py:113-122, LCUStep.py:66-76, SUMStep.py:37-47): the unitary of a circuit, column by
column.  Gate kinds follow include/qkan_b200.h (qkan_simulate_circuit)."""
import numpy as np

H, RY, CX, SWAP, X, Z = 0, 1, 2, 3, 4, 5


def evolve(gates, params, n_qubits, basis_state):
    S = 1 << n_qubits
    psi = np.zeros(S, dtype=np.complex128)
    psi[basis_state] = 1
    idx = np.arange(S)
    for (kind, q0, q1), th in zip(gates, params):
        if kind in (H, RY, X, Z):
            lo = idx[(idx >> q0) & 1 == 0]
            hi = lo | (1 << q0)
            a, b = psi[lo].copy(), psi[hi].copy()
            if kind == H:
                psi[lo], psi[hi] = (a + b) / np.sqrt(2), (a - b) / np.sqrt(2)
            elif kind == RY:
                c, s = np.cos(th / 2), np.sin(th / 2)
                psi[lo], psi[hi] = c * a - s * b, s * a + c * b
            elif kind == X:
                psi[lo], psi[hi] = b, a
            else:
                psi[hi] = -b
        elif kind == CX:
            lo = idx[((idx >> q1) & 1 == 0) & ((idx >> q0) & 1 == 1)]
            hi = lo | (1 << q1)
            psi[lo], psi[hi] = psi[hi].copy(), psi[lo].copy()
        elif kind == SWAP:
            a = idx[((idx >> q0) & 1 == 1) & ((idx >> q1) & 1 == 0)]
            b = (a ^ (1 << q0)) | (1 << q1)
            psi[a], psi[b] = psi[b].copy(), psi[a].copy()
        else:
            raise ValueError(kind)
    return psi


def top_left_block(gates, params, n_qubits, size):
    cols = np.stack([evolve(gates, params, n_qubits, j) for j in range(size)])
    return cols[:, :size].T
