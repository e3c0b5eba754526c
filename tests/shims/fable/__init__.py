"""`from fable import fable` of the reference (ChebyshevStep.py:3, MulStep.py:3, LCUStep.py:4, SUMStep.py:5) -> the
product's FABLE gate-list generator.  Test shim, see tests/shims/README.md."""
from qkan_implementation_b200.fable import fable  # noqa: F401
