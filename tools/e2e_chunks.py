"""Time the host-buffer forward (pinned) for different chunk counts.  GPU only."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer
N, K, D, B = 4, 4, 3, 1_000_000
x = (torch.rand((B, N), dtype=torch.float64) * 2 - 1).pin_memory()
o = torch.empty((B, K), dtype=torch.float64).pin_memory()
W = list((np.random.default_rng(0).uniform(-1, 1, (D + 1, N * K))))
layer = QKANLayer(N, K, D)
xn, on = x.numpy(), o.numpy()
for ch in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 64):
    os.environ["QKAN_HOST_CHUNKS"] = str(ch)
    for _ in range(3):
        layer.forward(xn, W, out=on, check_range=False)
    t0 = time.perf_counter()
    for _ in range(20):
        layer.forward(xn, W, out=on, check_range=False)
    dt = (time.perf_counter() - t0) / 20
    print(f"chunks {ch:3d}: {dt*1e3:.3f} ms  {B/dt/1e9:.3f} Gsamples/s  {(B*(N+K)*8)/dt/1e9:.1f} GB/s both ways")
# raw copies for reference
xd = torch.empty((B, N), dtype=torch.float64, device="cuda")
for _ in range(3): xd.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): xd.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"raw H2D 32 MB: {dt*1e3:.3f} ms {32e6/dt/1e9:.1f} GB/s")
