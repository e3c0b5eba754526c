"""N-GPU check of the sharded degree evaluation (one process per GPU, NCCL): every rank scores its slice of the rows,
the Gram matrices / residual sums are all-reduced, and the result must equal the single-GPU scores.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/degree_multi_gpu.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares  # noqa: E402
from qkan_implementation_b200.distributed import shard_bounds  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
n, F, D = 774_456, 79, 3
gen = torch.Generator().manual_seed(0)
x = torch.randn((n, F), dtype=torch.float64, generator=gen) * 0.6
y = torch.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.1 * torch.randn(n, dtype=torch.float64, generator=gen)
w = torch.rand(n, dtype=torch.float64, generator=gen) + 0.5
lo, hi = shard_bounds(n, world, rank)
xs, ys, ws = x[lo:hi].cuda(), y[lo:hi].cuda().contiguous(), w[lo:hi].cuda().contiguous()
eng = ChebyshevLeastSquares(D, group=dist.group.WORLD)
scores, r2 = eng.solve(xs, ys, ws)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(3):
    eng.solve(xs, ys, ws)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
if rank == 0:
    one = ChebyshevLeastSquares(D)
    s1, r1 = one.solve(x.cuda(), y.cuda().contiguous(), w.cuda().contiguous())
    print({"world": world, "rows_per_rank": hi - lo, "seconds_per_evaluate": dt, "scores": scores.tolist(),
           "max_rel_diff_vs_one_gpu": float(np.max(np.abs(scores - s1) / np.abs(s1))), "max_abs_diff_r2": float(np.max(np.abs(r2 - r1)))})
    assert np.max(np.abs(scores - s1) / np.abs(s1)) < 1e-10
dist.barrier()
dist.destroy_process_group()
