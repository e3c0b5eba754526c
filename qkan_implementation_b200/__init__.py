"""qkan_implementation_b200 - B200-native batched QKANLayer.forward.

Drop-in for the reference's ``QKAN_Steps_original`` package on the forward path:

    from qkan_implementation_b200 import QKANLayer
    layer = QKANLayer(N=4, K=4, max_degree=3)
    y = layer.forward(x, weights)            # x [N] or [B, N]; NumPy or torch (CPU / CUDA)

All arithmetic runs in ``libqkan_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/qkan_b200.h``); there is no CPU fallback.
"""
from .layer import ChebyshevStep, LCUStep, MulStep, QKANLayer, SUMStep
from . import _binding
from .distributed import FusedGatherQKANLayer, ShardedQKANLayer, shard_bounds
from .degree_optimizer import DegreeOptimizer

__all__ = ["ChebyshevStep", "MulStep", "LCUStep", "SUMStep", "QKANLayer", "ShardedQKANLayer", "FusedGatherQKANLayer", "shard_bounds", "DegreeOptimizer"]
__version__ = "0.1.0"
