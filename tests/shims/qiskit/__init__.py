"""Names the reference imports from qiskit (MulStep.py:4, LCUStep.py:5, SUMStep.py:6, ChebyshevStep.py:4).  Only
`transpile` is ever called, by the in-file tests: the gate list needs no compilation for the simulator behind the
qiskit_aer shim.  Test shim, see tests/shims/README.md."""


def transpile(circuit, backend=None, **_kwargs):
    return circuit


class QuantumCircuit:            # imported by the reference, never instantiated on the tested paths
    pass


class QuantumRegister:
    pass


class ClassicalRegister:
    pass
