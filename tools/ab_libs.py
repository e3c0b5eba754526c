#!/usr/bin/env python
"""A/B of several builds of libqkan_b200.so on one box: kernel-only time of a list of workloads per library.
    python tools/ab_libs.py --libs name=path,name=path [--configs c2,c3,c4,c5d1..c5d16] [--dtype complex128]
Every (library, workload) runs in its own process (the library path is read at import: QKAN_B200_LIB)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = {"c2": (4, 4, 3, 1_000_000), "c3": (16, 16, 8, 1_000_000), "c4": (784, 10, 5, 100_000), "c3s": (16, 16, 8, 200_000)}
for d in range(1, 17):
    CONFIGS[f"c5d{d}"] = (8, 8, d, 2_000_000)


def child(a):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from qkan_implementation_b200 import QKANLayer, _binding
    peak = _binding.measure_fma_peak(0, a.dtype != "complex64")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in a.configs.split(","):
        N, K, D, B = CONFIGS[name]
        gen = torch.Generator().manual_seed(0)
        x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
        W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
        layer = QKANLayer(N, K, D, dtype=a.dtype)
        y = layer.forward(x, W)
        ref = torch.cos(D * torch.acos(x.clamp(-1, 1)))
        # closed form of the compat-mode layer (oracle-free sanity check of the build under test)
        idx = (torch.arange(N * K, device="cuda") // K)
        wm = W.mean(0)
        cf = (ref[:, idx] * wm).reshape(B, K, N).sum(2) / N
        err = float((y - cf).abs().max())
        for _ in range(3):
            layer._engine.forward_device(x, False)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(a.reps):
            flush.zero_()
            ev0.record()
            layer._engine.forward_device(x, False)
            ev1.record()
            ev1.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        ms = float(np.median(ts))
        info = layer.kernel_info()
        print(json.dumps({"lib": a.tag, "cfg": name, "dtype": a.dtype, "ms": round(ms, 4), "samples_per_s": round(B / (ms * 1e-3)),
                          "frac": round(info["flops_exec"] * B / (ms * 1e-3) / 1e12 / peak, 3), "form": info["scaled_rotations"],
                          "flops_exec": info["flops_exec"], "max_abs_err_vs_closed_form": err, "peak": round(peak, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="")
    ap.add_argument("--configs", default="c2,c5d1,c5d2,c5d3,c5d4,c5d6,c5d8,c5d10,c5d12,c5d16,c3s,c4")
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    if a.tag:
        return child(a)
    for spec in a.libs.split(","):
        tag, path = spec.split("=")
        env = dict(os.environ, QKAN_B200_LIB=os.path.abspath(path))
        subprocess.run([sys.executable, __file__, "--tag", tag, "--configs", a.configs, "--dtype", a.dtype, "--reps", str(a.reps)], env=env, check=False)


if __name__ == "__main__":
    main()
