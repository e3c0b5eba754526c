"""Same-box A/B of launch-time tuning knobs (environment variables read by libqkan_b200.so at layer creation /
launch).  GPU only.
    python tools/ab_env.py c2 "" "QKAN_BLOCK_ROW_WORDS=36" "QKAN_BLOCK_SUB=2" "QKAN_BLOCK_TUNE=1:256:3:2"
Settings are measured interleaved, `--reps` rounds of `--iters` launches each; prints the median ms per setting."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import QKANLayer  # noqa: E402
from tools.tune import CONFIGS  # noqa: E402

KNOBS = ("QKAN_BLOCK_ROW_WORDS", "QKAN_BLOCK_SUB", "QKAN_BLOCK_STRIDED", "QKAN_BLOCK_TUNE", "QKAN_BLOCK_NO_DT")


def apply(setting):
    for k in KNOBS:
        os.environ.pop(k, None)
    for kv in setting.split():
        k, v = kv.split("=")
        os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("settings", nargs="+")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    N, K, D, B = CONFIGS[a.config]
    if a.batch:
        B = a.batch
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand((B, N), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
    W = (torch.rand((D + 1, N * K), dtype=torch.float64, generator=gen) * 2 - 1).cuda()
    layers, infos, times = [], [], [[] for _ in a.settings]
    for s in a.settings:
        apply(s)
        layer = QKANLayer(N, K, D)
        layer.forward(x, W)
        layers.append(layer)
        infos.append(layer.kernel_info())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.reps):
        for i, s in enumerate(a.settings):
            apply(s)
            for _ in range(3):
                layers[i]._engine.forward_device(x, False)
            for _ in range(a.iters):
                flush.zero_()
                ev0.record()
                layers[i]._engine.forward_device(x, False)
                ev1.record()
                ev1.synchronize()
                times[i].append(ev0.elapsed_time(ev1))
    for i, s in enumerate(a.settings):
        info = layers[i].kernel_info()
        ms = float(np.median(times[i]))
        print(json.dumps({"cfg": a.config, "B": B, "setting": s, "ms": round(ms, 4), "samples_per_s": round(B / ms * 1e3),
                          "NT": info["threads_per_cta"], "SU": info["samples_per_lane"], "MINB": info["min_ctas_per_sm"],
                          "grid": info["grid"], "smem": info["smem_bytes"]}))


if __name__ == "__main__":
    main()
