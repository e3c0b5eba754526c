"""Sensitivity of the block kernel to the x-tile size (env QKAN_BLOCK_SUB).  GPU only."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer
for (N, K, D, B) in ((4, 4, 3, 1_000_000), (8, 8, 1, 1_000_000), (8, 8, 4, 500_000)):
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
    W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
    for sub in (1, 2, 4, 8, 16, 32):
        os.environ["QKAN_BLOCK_SUB"] = str(sub)
        layer = QKANLayer(N, K, D)
        for _ in range(3):
            layer.forward(x, W)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(10):
            e0.record(); layer._engine.forward_device(x, False); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        info = layer.kernel_info()
        print((N, K, D), "sub", sub, "grid", info["grid"], "smem", info["smem_bytes"], "ms %.4f" % np.median(ts), "Msps %.0f" % (B / np.median(ts) / 1e3))
