#!/usr/bin/env python
"""Where the end-to-end time of the pinned-host path goes: the forward kernel timed with CUDA events on
(x host | device) x (out host | device), against the wall clock of the public call.   python tools/e2e_breakdown.py [N K D B]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer, _binding as b  # noqa: E402

N, K, D, B = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 4, 3, 1_000_000)
gen = torch.Generator().manual_seed(0)
x = torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1
W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).numpy()
xh = torch.empty((B, N), dtype=torch.float64).pin_memory()
xh.copy_(x)
oh = torch.empty((B, K), dtype=torch.float64).pin_memory()
xd, od = x.cuda(), torch.empty((B, K), dtype=torch.float64, device="cuda")
layer = QKANLayer(N, K, D)
layer.forward(xd, W)
lib, h = b.lib(), layer._engine.handle()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, xi, oi in (("x device, out device", xd, od), ("x host,   out device", xh, od), ("x device, out host  ", xd, oh), ("x host,   out host  ", xh, oh)):
    ts = []
    for it in range(25):
        e0.record()
        b.check(lib.qkan_layer_forward(h, xi.data_ptr(), B, oi.data_ptr(), None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[5:]))
    gb = ((N * 8 * B) if xi is xh else 0) + ((K * 8 * B) if oi is oh else 0)
    print(f"kernel, {name}: {ms:.3f} ms" + (f"  ({gb / ms / 1e6:.1f} GB/s over PCIe, both directions summed)" if gb else ""))
xn, on = xh.numpy(), oh.numpy()
Wl = list(W)
for name, kw in (("forward(x, W, out=)                 ", {}), ("forward(x, W, out=, check_range=False)", {"check_range": False})):
    for _ in range(3):
        layer.forward(xn, Wl, out=on, **kw)
    ts = []
    for _ in range(25):
        t0 = time.perf_counter()
        layer.forward(xn, Wl, out=on, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"public call {name}: {float(np.median(ts)) * 1e3:.3f} ms wall")
ts = []
for _ in range(25):
    t0 = time.perf_counter()
    b.check(lib.qkan_layer_forward_host(h, xn.ctypes.data, B, on.ctypes.data, None))
    ts.append(time.perf_counter() - t0)
print(f"C ABI qkan_layer_forward_host: {float(np.median(ts)) * 1e3:.3f} ms wall")
