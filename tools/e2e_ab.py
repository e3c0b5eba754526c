"""Host-buffer path A/B: pinned zero-copy (kernel reads / writes host memory) versus the staged copy pipeline.
    python tools/e2e_ab.py [N K D B]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer  # noqa: E402

N, K, D, B = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 4, 3, 1_000_000)
gen = torch.Generator().manual_seed(0)
x = torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1
W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).numpy()
xh = torch.empty((B, N), dtype=torch.float64).pin_memory()
xh.copy_(x)
oh = torch.empty((B, K), dtype=torch.float64).pin_memory()
xn, on = xh.numpy(), oh.numpy()
layer = QKANLayer(N, K, D)
ref = layer.forward(x.cuda(), W).cpu().numpy()
for mode, chunks, graph, edge in (("zero_copy", 0, 1, 4), ("staged", 4, 0, 1), ("staged", 8, 1, 1), ("staged", 0, 1, 1), ("staged", 0, 1, 4), ("staged", 32, 1, 1),
                                  ("copy_in", 4, 0, 1), ("copy_in", 8, 1, 1), ("copy_in", 8, 1, 4), ("copy_in", 0, 1, 1), ("copy_in", 0, 1, 2), ("copy_in", 0, 1, 4),
                                  ("copy_in", 0, 1, 8), ("copy_in", 16, 1, 4), ("copy_in", 0, 0, 4), ("copy_out", 0, 1, 4), ("zero_copy", 0, 1, 4)):
    os.environ["QKAN_HOST_EDGE_DIV"] = str(edge)
    if graph:
        os.environ.pop("QKAN_HOST_NO_GRAPH", None)
    else:
        os.environ["QKAN_HOST_NO_GRAPH"] = "1"
    os.environ["QKAN_HOST_PATH"] = mode
    if chunks:
        os.environ["QKAN_HOST_CHUNKS"] = str(chunks)
    else:
        os.environ.pop("QKAN_HOST_CHUNKS", None)
    on[:] = 0
    for _ in range(3):
        layer.forward(xn, W, out=on, check_range=False)
    assert np.array_equal(on, ref), mode
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        layer.forward(xn, W, out=on, check_range=False)
        ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    print(f"{mode:10s} chunks={chunks or 'auto':>4} edge_div={edge} graph={graph} N={N} K={K} D={D} B={B}: {ms:.3f} ms  {B / ms / 1e6:.3f} Gsamples/s  {(N + K) * 8 * B / ms / 1e6:.1f} GB/s both ways")
os.environ.pop("QKAN_HOST_PATH", None)
os.environ.pop("QKAN_HOST_EDGE_DIV", None)
os.environ.pop("QKAN_HOST_CHUNKS", None)
os.environ.pop("QKAN_HOST_NO_GRAPH", None)
# pageable numpy buffers always take the staged path
xp, op = x.numpy().copy(), np.empty((B, K))
layer.forward(xp, W, out=op, check_range=False)
assert np.array_equal(op, ref)
print("pageable buffers ok (staged)")
