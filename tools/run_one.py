"""Run one workload a few times (for ncu / compute-sanitizer).  python tools/run_one.py N K D B [dtype] [prep]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer  # noqa: E402

N, K, D, B = map(int, sys.argv[1:5])
dtype = sys.argv[5] if len(sys.argv) > 5 else "complex128"
prep = sys.argv[6] if len(sys.argv) > 6 else "analytic"
gen = torch.Generator().manual_seed(0)
x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
layer = QKANLayer(N, K, D, dtype=dtype, prep=prep)
for _ in range(4):
    y = layer.forward(x, W)
torch.cuda.synchronize()
print(layer.kernel_info(), float(y.abs().max()))
