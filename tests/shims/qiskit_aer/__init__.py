"""`Aer.get_backend('unitary_simulator')` of the reference's tests (MulStep.py:113, LCUStep.py:66, SUMStep.py:37,
ChebyshevStep.py:125): run(circuit).result().get_unitary(circuit) returns the full 2^(2n+1)-square unitary of a
FableCircuit.  GPU present: qkan_simulate_circuit (the product's batched gate-list simulator) evolves every basis state;
no GPU: oracle/circuit_sim.py.  Qiskit's convention is kept: qubit q is bit q of the row / column index.
Test shim, see tests/shims/README.md."""
import numpy as np

BACKEND_USED = []        # "gpu" / "oracle", appended per run (the tests report which simulator checked the circuits)


def _unitary(circuit):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:      # noqa: BLE001
        have_gpu = False
    if have_gpu:
        BACKEND_USED.append("gpu")
        cols = circuit.columns(np.arange(1 << circuit.num_qubits))      # row j = U |j>
        return np.ascontiguousarray(cols.T)
    from oracle import circuit_sim
    BACKEND_USED.append("oracle")
    return circuit_sim.unitary(circuit.gates, circuit.params, circuit.num_qubits)


class _Result:
    def __init__(self, circuit):
        self._circuit = circuit
        self._u = None

    def get_unitary(self, circuit=None, decimals=None):
        if self._u is None:
            self._u = _unitary(self._circuit if circuit is None else circuit)
        return self._u


class _Job:
    def __init__(self, circuit):
        self._circuit = circuit

    def result(self):
        return _Result(self._circuit)


class _UnitarySimulator:
    name = "unitary_simulator"

    def run(self, circuit, **_kwargs):
        return _Job(circuit)


class _Aer:
    @staticmethod
    def get_backend(name):
        if name != "unitary_simulator":
            raise ValueError(f"the test shim only provides 'unitary_simulator', not {name!r}")
        return _UnitarySimulator()


Aer = _Aer()
AerSimulator = _UnitarySimulator
