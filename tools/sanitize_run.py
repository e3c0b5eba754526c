"""Small forwards through every kernel family, ragged batches, edge inputs (for compute-sanitizer).
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer  # noqa: E402
from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares  # noqa: E402

rng = np.random.default_rng(0)
cases = [  # (N, K, D, B, kwargs)   direct (compile-time G and run-time G), element-owner (both loop orders), tile, generic, paper, gates
    (4, 4, 3, 1001, {}), (8, 8, 1, 777, {}), (16, 16, 8, 333, {}), (2, 2, 2, 130, {}), (4, 8, 2, 515, {}),
    (784, 10, 5, 259, {}), (100, 10, 5, 1300, {}), (33, 3, 2, 401, {}),
    (3, 5, 2, 600, {}), (5, 3, 1, 77, {}), (7, 1, 3, 99, {}),
    (4, 4, 20, 200, {}), (4, 4, 0, 50, {}), (4, 4, 3, 300, {"mode": "paper"}), (4, 4, 3, 257, {"prep": "gates"}),
]
for (N, K, D, B, kw) in cases:
    for dtype in ("complex128", "complex64"):
        if kw.get("prep") == "gates" and dtype != "complex128":
            continue
        x = rng.uniform(-1, 1, (B, N))
        x[0, 0] = 1.0
        x[B - 1, N - 1] = -1.5
        W = rng.uniform(-1, 1, (D + 1, N * K))
        for order in ("0", "1"):
            os.environ["QKAN_ELEM_ROW_OUTER"] = order
            layer = QKANLayer(N, K, D, dtype=dtype, **kw)
            y = layer.forward(torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda())
            y2, amps = layer.forward(torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda(), return_amplitudes=True)
            yh = layer.forward(x, W)                          # host path (pageable buffers: staged chunks)
            torch.cuda.synchronize()
            assert np.allclose(yh, y.cpu().numpy(), atol=1e-4)
            if not layer.kernel_info()["element_owner"]:
                break
    print("ok", N, K, D, B, kw, flush=True)
os.environ.pop("QKAN_ELEM_ROW_OUTER", None)
# pinned host buffers (chunk pipeline, kernel-written results)
xh = torch.from_numpy(rng.uniform(-1, 1, (300_001, 4))).pin_memory()
oh = torch.empty((300_001, 4), dtype=torch.float64).pin_memory()
layer = QKANLayer(4, 4, 3)
layer.forward(xh.numpy(), rng.uniform(-1, 1, (4, 16)), out=oh.numpy())
layer.forward(xh.numpy(), rng.uniform(-1, 1, (4, 16)), out=oh.numpy())
print("ok pinned host path", flush=True)
# degree evaluation kernels (Gram + reduce + residuals), compile-time and run-time degrees, ragged row counts
for (n, F, D) in ((5003, 7, 3), (4099, 79, 1), (3001, 5, 6)):
    xs = torch.from_numpy(rng.uniform(-1.2, 1.2, (n, F))).cuda()
    ys = torch.from_numpy(rng.normal(size=n)).cuda()
    eng = ChebyshevLeastSquares(D)
    G = eng.gram(xs, ys)
    torch.cuda.synchronize()
    print("ok gram", n, F, D, float(G[0, 0]), flush=True)
print("done")
