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


def unitary(gates, params, n_qubits, columns=None):
    """Columns `columns` (default: all) of the circuit's unitary, [2^n_qubits, len(columns)] complex128.

    A maximal run of consecutive gates with the same target qubit t (H / RY / X / Z on t, CX with target t) is folded
    into one 2 x 2 matrix per value of the other qubits before it touches the state, so FABLE's oracle (2 * 4^n gates,
    all on the flag qubit) costs one pass - the same run fusion as csrc/qkan_circuit.cu, restated independently."""
    S = 1 << n_qubits
    cols = np.arange(S) if columns is None else np.asarray(columns)
    psi = np.zeros((S, len(cols)), dtype=np.complex128)
    psi[cols, np.arange(len(cols))] = 1
    idx = np.arange(S)
    gates = [tuple(g) for g in gates]
    g = 0
    while g < len(gates):
        kind, q0, q1 = gates[g]
        if kind == SWAP:
            a = idx[((idx >> q0) & 1 == 1) & ((idx >> q1) & 1 == 0)]
            b = (a ^ (1 << q0)) | (1 << q1)
            psi[a], psi[b] = psi[b].copy(), psi[a].copy()
            g += 1
            continue
        t = q1 if kind == CX else q0
        lo = idx[(idx >> t) & 1 == 0]
        hi = lo | (1 << t)
        m00 = np.ones(len(lo)); m01 = np.zeros(len(lo)); m10 = np.zeros(len(lo)); m11 = np.ones(len(lo))
        while g < len(gates):
            k, a0, a1 = gates[g]
            if k == SWAP or (a1 if k == CX else a0) != t:
                break
            if k == RY:
                c, s = np.cos(params[g] / 2), np.sin(params[g] / 2)
                m00, m01, m10, m11 = c * m00 - s * m10, c * m01 - s * m11, s * m00 + c * m10, s * m01 + c * m11
            elif k == H:
                r = 1 / np.sqrt(2)
                m00, m01, m10, m11 = (m00 + m10) * r, (m01 + m11) * r, (m00 - m10) * r, (m01 - m11) * r
            elif k == X:
                m00, m01, m10, m11 = m10, m11, m00, m01
            elif k == Z:
                m10, m11 = -m10, -m11
            elif k == CX:
                on = ((lo >> a0) & 1) == 1
                m00, m10 = np.where(on, m10, m00), np.where(on, m00, m10)
                m01, m11 = np.where(on, m11, m01), np.where(on, m01, m11)
            else:
                raise ValueError(k)
            g += 1
        a, b = psi[lo].copy(), psi[hi].copy()
        psi[lo] = m00[:, None] * a + m01[:, None] * b
        psi[hi] = m10[:, None] * a + m11[:, None] * b
    return psi
