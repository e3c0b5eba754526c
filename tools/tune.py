"""Time every tuning variant (env QKAN_VARIANT) of the kernel for a few workloads.  GPU only.
    python tools/tune.py [--configs c2,c5d1,c5d4,c5d16,c3] [--variants 0,1,2,...]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer, _binding  # noqa: E402

CONFIGS = {"c4b": (784, 10, 5, 100_000), "c2": (4, 4, 3, 1_000_000), "c5d1": (8, 8, 1, 1_000_000), "c5d2": (8, 8, 2, 1_000_000), "c5d4": (8, 8, 4, 500_000),
           "c5d8": (8, 8, 8, 200_000), "c5d16": (8, 8, 16, 100_000), "c3": (16, 16, 8, 50_000), "c4": (784, 10, 5, 20_000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c5d1,c5d4,c5d16,c3")
    ap.add_argument("--variants", default="0:0:0:0,1:256:4:1,1:256:3:2,1:256:2:4,1:128:8:1",
                    help="QKAN_BLOCK_TUNE values U:NT:MINB:SU (0 = planner default); only built combinations resolve")
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--prep", default="analytic")
    a = ap.parse_args()
    peak = _binding.measure_fma_peak(0, a.dtype != "complex64")
    print(f"DFMA peak {peak:.2f} TFLOP/s")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in a.configs.split(","):
        N, K, D, B = CONFIGS[name]
        gen = torch.Generator().manual_seed(0)
        x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
        W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
        seen = set()
        ref = None
        for v in a.variants.split(","):
            os.environ["QKAN_BLOCK_TUNE" if a.prep == "analytic" else "QKAN_VARIANT"] = v
            try:
                layer = QKANLayer(N, K, D, dtype=a.dtype, prep=a.prep)
                y = layer.forward(x, W)
            except Exception as e:   # noqa
                print(name, "variant", v, "failed:", e)
                continue
            info = layer.kernel_info()
            key = (info["unroll"], info["samples_per_lane"], info["lanes_per_sample"], info["tile_qubits"], info["local_qubits"], info["threads_per_cta"], info["min_ctas_per_sm"], info["grid"], info["smem_bytes"])
            if key in seen:
                continue
            seen.add(key)
            if ref is None:
                ref = y
            same = bool(torch.equal(y, ref))
            for _ in range(3):
                layer.forward(x, W)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(10):
                flush.zero_()                              # evict x / out from L2; also lets the host run ahead of the GPU
                ev0.record()
                layer._engine.forward_device(x, False)
                ev1.record()
                ev1.synchronize()
                ts.append(ev0.elapsed_time(ev1))
            ms = float(np.median(ts))
            tf = info["flops_exec"] * B / (ms * 1e-3) / 1e12
            print(json.dumps({"cfg": name, "variant": v, "U": info["unroll"], "SU": info["samples_per_lane"], "G": info["lanes_per_sample"], "NT": info["threads_per_cta"],
                              "MINB": info["min_ctas_per_sm"], "passes": info["passes"], "rows": info["row_steps"],
                              "tileq": info["tile_qubits"], "T": info["local_qubits"], "grid": info["grid"],
                              "cta_per_sm": round(info["grid"] / 148, 2), "smem": info["smem_bytes"], "ms": round(ms, 4),
                              "samples_per_s": round(B / (ms * 1e-3)), "tflops_exec": round(tf, 2),
                              "frac": round(tf / peak, 3),
                              "frac_per_block_basis": round(info["flops_per_block_basis"] * B / (ms * 1e-3) / 1e12 / peak, 3),
                              "pipe_util": round(info["fp_inst_exec"] * B / (ms * 1e-3) / (peak * 1e12 / 2), 3),
                              "frac_survey": round(info["flops_survey"] * B / (ms * 1e-3) / 1e12 / peak, 3), "bitwise_same": same}))
    os.environ.pop("QKAN_VARIANT", None)
    os.environ.pop("QKAN_BLOCK_TUNE", None)


if __name__ == "__main__":
    main()
