"""Host path experiment: explicit chunk boundaries (QKAN_HOST_CUTS).   python tools/e2e_cuts.py [N K D B]"""
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
print("x addr % 2MiB", xh.data_ptr() % (2 << 20), "out addr % 2MiB", oh.data_ptr() % (2 << 20))
layer = QKANLayer(N, K, D)
ref = layer.forward(x.cuda(), W).cpu().numpy()


def uniform(step):
    return list(range(step, B, step))


def ramp(first, full):
    cuts, pos, c = [], 0, first
    while pos + c < B:
        pos += c
        cuts.append(pos)
        c = min(full, c * 2)
    # ramp down at the end: split the last stretch
    return cuts


u = 8 * 1024 * 1024 // ((N + K) * 8)     # samples per 8 MiB of traffic
cases = {"auto": None, "uniform u": uniform(u), "uniform u/2": uniform(u // 2), "uniform 2u": uniform(2 * u), "uniform 125056": uniform(125056),
         "ramp u/8..u": ramp(u // 8, u), "ramp u/4..u": ramp(u // 4, u), "ramp u/4..2u": ramp(u // 4, 2 * u),
         "ramp u/4..u + tail": None, "auto again": None}
t = ramp(u // 4, u)
tail = [B - u // 4 - u // 2, B - u // 4]
cases["ramp u/4..u + tail"] = [c for c in t if c < tail[0] - u // 2] + tail
for name, cuts in cases.items():
    if cuts is None:
        os.environ.pop("QKAN_HOST_CUTS", None)
    else:
        os.environ["QKAN_HOST_CUTS"] = ",".join(str(c) for c in cuts)
    on[:] = 0
    for _ in range(3):
        layer.forward(xn, W, out=on, check_range=False)
    assert np.array_equal(on, ref), name
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        layer.forward(xn, W, out=on, check_range=False)
        ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    print(f"{name:22s} chunks={len(cuts) + 1 if cuts is not None else 'auto':>4} N={N} K={K} D={D} B={B}: {ms:.3f} ms  {B / ms / 1e6:.3f} Gsamples/s  {(N + K) * 8 * B / ms / 1e6:.1f} GB/s both ways", flush=True)
