"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): the kernel-fused output
gather - NVLS multicast and NVLink peer stores - and the NCCL gather, each compared BITWISE with the single-GPU result of
the whole batch, on ragged batches and for the direct, tile and window kernels.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import FusedGatherQKANLayer, QKANLayer, ShardedQKANLayer  # noqa: E402
from qkan_implementation_b200.distributed import shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = []
    paths = set()
    for (N, K, D, B) in ((4, 4, 3, 100_003), (8, 8, 1, 50_001), (8, 8, 16, 20_011), (5, 3, 2, 33_333), (100, 10, 5, 4_097), (16, 16, 8, 7)):
        rng = np.random.default_rng(N + K + D)                # the same inputs on every rank
        x = torch.from_numpy(rng.uniform(-1, 1, (B, N))).to(dev)
        W = list(rng.uniform(-1, 1, (D + 1, N * K)))
        layer = QKANLayer(N, K, D, device=local)
        ref = layer.forward(x, W)                             # the whole batch on this GPU
        lo, hi = shard_bounds(B, world, rank)
        for mc in (True, False):
            fused = FusedGatherQKANLayer(layer, multicast=mc)
            for rep in range(3):                              # repeated calls alternate the two result buffers
                y = fused.forward(x[lo:hi].contiguous(), W, B)
                paths.add(fused.last_path)
                if not torch.equal(y, ref):
                    failures.append(f"fused {fused.last_path} N{N} K{K} D{D} B{B} rep {rep}: max diff {float((y - ref).abs().max())}")
            dist.barrier()
        y = ShardedQKANLayer(layer).forward(x, W)             # NCCL all-gather of the sharded outputs
        if not torch.equal(y, ref):
            failures.append(f"nccl gather N{N} K{K} D{D} B{B}")
    bad = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(bad)
    if rank == 0:
        print(f"world {world}: fused-gather paths exercised {sorted(paths)}; failures on all ranks: {int(bad.item())}")
    for f in failures:
        print(f"rank {rank}: {f}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(bad.item()) else 0)


if __name__ == "__main__":
    main()
