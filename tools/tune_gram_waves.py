"""Gram step (qkan_cheb_gram: kernel + reduce + expand) against the number of CTA waves the samples are sliced into
(QKAN_GRAM_WAVES, read per call).  GPU only.   python tools/tune_gram_waves.py [--n 774456 --F 79 --D 3]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=774_456)
    ap.add_argument("--F", type=int, default=79)
    ap.add_argument("--D", type=int, default=3)
    ap.add_argument("--waves", type=str, default="2,3,4,5,6,8,12")
    a = ap.parse_args()
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((a.n, a.F), dtype=torch.float64, device="cuda", generator=gen) * 0.6
    y = torch.randn(a.n, dtype=torch.float64, device="cuda", generator=gen)
    eng = ChebyshevLeastSquares(a.D)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    for wv in [int(v) for v in a.waves.split(",")]:
        os.environ["QKAN_GRAM_WAVES"] = str(wv)
        ts = []
        for _ in range(8):
            flush.zero_()
            ev0.record()
            G = eng.gram(x, y)
            ev1.record()
            ev1.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        if ref is None:
            ref = G.clone()
        print(json.dumps({"waves": wv, "n": a.n, "F": a.F, "D": a.D, "ms": round(float(np.median(ts[2:])), 4),
                          "max_rel_diff_vs_first": float((G - ref).abs().max() / ref.abs().max())}), flush=True)


if __name__ == "__main__":
    main()
