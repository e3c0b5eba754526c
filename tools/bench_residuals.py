"""Time qkan_cheb_residuals alone (CUDA events on the launching stream, L2 flushed between launches) at the reference's
degree-evaluation shape, for both kernels (QKAN_RES_KERNEL=warp forces the warp-per-sample kernel).  GPU only.
    python tools/bench_residuals.py [--n 774456 --F 79 --D 3]
One JSON line per (kernel, want_xtr): ms per launch, GB/s of x read, FP64 instruction count basis."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import _binding as _b  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=774_456)
    ap.add_argument("--F", type=int, default=79)
    ap.add_argument("--D", type=int, default=3)
    a = ap.parse_args()
    n, F, D = a.n, a.F, a.D
    D1, P = D + 1, F * (D + 1)
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, F), dtype=torch.float64, device="cuda", generator=gen) * 0.6
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    w = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    coef = torch.randn((D1, P), dtype=torch.float64, device="cuda", generator=gen) * 0.1
    lib = _b.lib()
    ctas = ctypes.c_int()
    _b.check(lib.qkan_cheb_residuals_ctas(ctypes.byref(ctas)))
    c = ctas.value
    sums = torch.empty((c, D1, 2), dtype=torch.float64, device="cuda")
    tail = torch.empty((c, 4), dtype=torch.float64, device="cuda")
    xtr = torch.empty((c, D1, P), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    results = {}
    for kernel in ("tile", "tile_64_rows", "warp"):
        os.environ.pop("QKAN_RES_KERNEL", None)
        os.environ.pop("QKAN_RES_SPL", None)
        if kernel == "warp":
            os.environ["QKAN_RES_KERNEL"] = "warp"
        elif kernel != "tile":
            os.environ["QKAN_RES_SPL"] = "1"
        for want_xtr in (True, False):
            ts = []
            for _ in range(9):
                flush.zero_()
                ev0.record()
                _b.check(lib.qkan_cheb_residuals(x.data_ptr(), y.data_ptr(), w.data_ptr(), n, F, D, coef.data_ptr(), 0.0,
                                                 sums.data_ptr(), tail.data_ptr(), xtr.data_ptr() if want_xtr else None, stream))
                ev1.record()
                ev1.synchronize()
                ts.append(ev0.elapsed_time(ev1))
            ms = float(np.median(ts[3:]))
            results[(kernel, want_xtr)] = (sums.sum(0).cpu().numpy(), xtr.sum(0).cpu().numpy() if want_xtr else None)
            print(json.dumps({"kernel": kernel, "want_xtr": want_xtr, "n": n, "F": F, "D": D, "ms": round(ms, 4),
                              "x_GBps": round(n * F * 8 / ms / 1e6, 1)}), flush=True)
    for want_xtr in (True, False):
        s_t, x_t = results[("tile", want_xtr)]
        s_w, x_w = results[("warp", want_xtr)]
        d = {"want_xtr": want_xtr, "max_rel_diff_sums_tile_vs_warp": float(np.abs(s_t - s_w).max() / np.abs(s_w).max())}
        if want_xtr:
            d["max_rel_diff_xtr_tile_vs_warp"] = float(np.abs(x_t - x_w).max() / np.abs(x_w).max())
        print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()
