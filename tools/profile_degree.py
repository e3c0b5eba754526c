"""Where the wall time of ChebyshevLeastSquares.solve goes (774 456 x 79, degree 3): the stages of solve() run one by one with
a device synchronisation and a host timer after each (so the sum is larger than the un-instrumented call, which is printed
beside it).  GPU only.   python tools/profile_degree.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares, _NestedSolver, _blas_single_thread  # noqa: E402


def main():
    n, F, D = 774_456, 79, 3
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, F), dtype=torch.float64, device="cuda", generator=gen) * 0.6
    y = (torch.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)).contiguous()
    w = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    eng = ChebyshevLeastSquares(D)
    for _ in range(3):
        eng.solve(x, y, w)
    torch.cuda.synchronize()
    whole = []
    for _ in range(7):
        t0 = time.perf_counter()
        eng.solve(x, y, w)
        torch.cuda.synchronize()
        whole.append(time.perf_counter() - t0)
    D1, P = D + 1, F * (D + 1)
    acc = {}

    def lap(name, t0):
        torch.cuda.synchronize()
        acc.setdefault(name, []).append(time.perf_counter() - t0)
        return time.perf_counter()

    for _ in range(7):
        t = time.perf_counter()
        Gd = eng.gram(x, y)
        t = lap("gram kernels", t)
        G = Gd.cpu().numpy()
        t = lap("G to host", t)
        ybar = G[0, P] / n
        with _blas_single_thread():
            solver = _NestedSolver(G, n, F, D)
        t = lap("Cholesky (host)", t)
        coef = np.zeros((D1, P))
        with _blas_single_thread():
            for d in range(D1):
                coef[d, :F * (d + 1)] = solver.solve(d, G[:F * (d + 1), P])
        t = lap("4 solves (host)", t)
        _, _, xr = eng.residual_sums(x, y, w, coef, ybar, True)
        t = lap("residual pass with X^T r (upload, kernel, reduce, download)", t)
        with _blas_single_thread():
            for d in range(D1):
                coef[d, :F * (d + 1)] += solver.solve(d, xr[d, :F * (d + 1)])
        t = lap("4 refinement solves (host)", t)
        eng.residual_sums(x, y, w, coef, ybar, False)
        t = lap("final residual pass", t)
    out = {"whole_call_ms": round(1e3 * float(np.median(whole)), 3),
           "stages_ms": {k: round(1e3 * float(np.median(v[2:])), 3) for k, v in acc.items()}}
    out["stages_sum_ms"] = round(sum(out["stages_ms"].values()), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
