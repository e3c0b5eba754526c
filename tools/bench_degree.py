"""Measure the degree-evaluation path at the reference's published workload shape:
774 456 rows x 79 features, max_degree 3 (README / output_result_*.txt of the reference).  GPU only.
    python tools/bench_degree.py [--n 774456 --F 79 --D 3 --cpu-rows 60000]
Prints one JSON line: evaluate_degree wall time through the drop-in DegreeOptimizer (device-resident data), the
Gram kernel's own time and FP64 roofline fraction, and the CPU oracle (the reference's algorithm: NumPy lstsq per
degree) timed on a bounded sample of the rows."""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import DegreeOptimizer, _binding  # noqa: E402
from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=774_456)
    ap.add_argument("--F", type=int, default=79)
    ap.add_argument("--D", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=60_000)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    n, F, D = a.n, a.F, a.D
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, F), dtype=torch.float64, device="cuda", generator=gen) * 0.6
    y = (torch.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)).contiguous()
    w = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    opt = DegreeOptimizer([F, 1], D)
    with contextlib.redirect_stdout(io.StringIO()):
        scores, r2 = opt.evaluate_degree(x, y, w)            # warm-up
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            opt.evaluate_degree(x, y, w)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    eng = ChebyshevLeastSquares(D)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    kt = []
    for _ in range(8):
        flush.zero_()
        ev0.record()
        eng.gram(x, y)
        ev1.record()
        ev1.synchronize()
        kt.append(ev0.elapsed_time(ev1))
    k_ms = float(np.median(kt[2:]))
    # flop bases (2 flops per multiply-add, symmetric half of A^T A).  The kernel keeps the F identical all-ones T_0 columns of
    # the design matrix ONCE (qkan_degree.cu, GramParams): W = F D + 2 distinct columns instead of P1 = F (D+1) + 1.
    P1 = F * (D + 1) + 1
    W = F * D + 2 if D >= 1 else P1
    T = (W + 63) // 64
    flops_alg = float(n) * W * (W + 1)                       # the distinct entries: what has to be computed
    flops_full = float(n) * P1 * (P1 + 1)                    # every column of the reference's design matrix (round 2's earlier basis)
    flops_exec = float(n) * (T * (T + 1) // 2) * 64 * 64 * 2  # whole 64 x 64 tiles of the upper triangle
    peak_fma = _binding.measure_fma_peak(0, True)
    peak = _binding.measure_dmma_peak(0)
    line = {"metric": "DegreeOptimizer.evaluate_degree wall time", "value": float(np.median(ts)), "unit": "s", "higher_is_better": False,
            "config": {"workload": f"{n} rows x {F} features, max_degree {D} ({D + 1} least-squares fits of up to {P1 - 1} columns), weighted metrics",
                       "data": "synthetic, device resident"},
            "scores": [float(v) for v in scores],
            "roofline": {"bound": "tensor", "kernel": "qkan_cheb_gram_kernel (+ reduce + expand)", "kernel_ms": k_ms,
                         "achieved": flops_alg / (k_ms * 1e-3) / 1e12, "achieved_executed": flops_exec / (k_ms * 1e-3) / 1e12,
                         "peak": peak, "unit": "TFLOP/s", "frac": flops_alg / (k_ms * 1e-3) / 1e12 / peak,
                         "frac_executed": flops_exec / (k_ms * 1e-3) / 1e12 / peak,
                         "frac_full_columns_basis": flops_full / (k_ms * 1e-3) / 1e12 / peak,
                         "algorithmic_flops": flops_alg, "full_columns_flops": flops_full, "distinct_columns": W, "full_columns": P1,
                         "algorithmic_bytes": float(n) * (F + 1) * 8,
                         "peak_source": "qkan_measure_dmma_peak: independent mma.sync.m8n8k4.f64 chains on all SMs, measured in this run",
                         "dfma_peak": peak_fma}}
    if not a.no_cpu:
        from oracle import degree_oracle as do
        m = min(a.cpu_rows, n)
        xs, ys, ws = x[:m].cpu().numpy(), y[:m].cpu().numpy(), w[:m].cpu().numpy()
        t0 = time.perf_counter()
        cs, _ = do.evaluate_degree(xs, ys, D, ws)
        dt = time.perf_counter() - t0
        with contextlib.redirect_stdout(io.StringIO()):
            gs, _ = opt.evaluate_degree(xs, ys, ws)
        line["cpu_baseline"] = {"value": dt * n / m, "unit": "s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"oracle/degree_oracle.evaluate_degree (NumPy lstsq per degree, threaded LAPACK) on the first {m} rows: "
                                          f"{dt:.2f} s, scaled linearly to {n} rows",
                                "max_rel_diff_scores_on_sample": float(np.max(np.abs(gs - cs) / np.abs(cs)))}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
