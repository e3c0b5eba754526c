"""Host path experiment: write-combined pinned input buffer (cudaHostAllocWriteCombined: not snooped during PCIe transfers)
against plain pinned memory.   python tools/e2e_wc.py [N K D B]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer  # noqa: E402

N, K, D, B = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 4, 3, 1_000_000)
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]


def host_alloc(shape, flags):
    n = int(np.prod(shape)) * 8
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), n, flags)
    assert rc == 0, rc
    buf = (ctypes.c_double * (n // 8)).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64).reshape(shape)


gen = torch.Generator().manual_seed(0)
x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1).numpy()
W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).numpy()
layer = QKANLayer(N, K, D)
ref = layer.forward(torch.from_numpy(x).cuda(), W).cpu().numpy()
MAPPED, WC = 2, 4
for name, fx, fo in (("pinned x, pinned out", MAPPED, MAPPED), ("write-combined x, pinned out", MAPPED | WC, MAPPED),
                     ("write-combined x, write-combined out", MAPPED | WC, MAPPED | WC), ("pinned x, pinned out (again)", MAPPED, MAPPED)):
    xn = host_alloc((B, N), fx)
    on = host_alloc((B, K), fo)
    xn[:] = x
    for path in ("copy_in", "zero_copy", "staged"):
        os.environ["QKAN_HOST_PATH"] = path
        for _ in range(3):
            layer.forward(xn, W, out=on, check_range=False)
        ok = np.array_equal(on, ref)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            layer.forward(xn, W, out=on, check_range=False)
            ts.append(time.perf_counter() - t0)
        ms = float(np.median(ts)) * 1e3
        t0 = time.perf_counter()
        s = float(on.sum())
        rd = (time.perf_counter() - t0) * 1e3
        print(f"{name:40s} {path:9s}: {ms:.3f} ms  {B / ms / 1e6:.3f} Gsamples/s  {(N + K) * 8 * B / ms / 1e6:.1f} GB/s both ways  equal={ok}  CPU sum of the result: {rd:.2f} ms", flush=True)
os.environ.pop("QKAN_HOST_PATH", None)
