#!/usr/bin/env python
"""Host<->device link ceiling for the e2e path: pinned-buffer copies with the DMA engines, one direction at a time and both
at once, per rank (run it under torchrun to load every GPU's link and the host memory at the same time).

    python tools/pcie_probe.py [MiB per direction, default 32]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import os
import sys
import time

import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("PROBE_BIND", "1") == "1":
    from bench import bind_to_gpu_numa_node
    bound = bind_to_gpu_numa_node(local)
else:
    bound = None
n = mib << 20
hin = torch.empty(n, dtype=torch.uint8).pin_memory()
hout = torch.empty(n, dtype=torch.uint8).pin_memory()
hin.fill_(1)
din = torch.empty(n, dtype=torch.uint8, device="cuda")
dout = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(h2d, d2h, reps=20):
    for it in range(3 + reps):
        if it == 3:
            barrier()
            t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9


res = {"h2d_only": run(True, False), "d2h_only": run(False, True), "both_each_way": run(True, True)}
line = f"rank {rank}/{world} gpu {local} bound_cpus {bound} {mib} MiB: " + "  ".join(f"{k} {v:.1f} GB/s" for k, v in res.items())
if world > 1:
    lines = [None] * world
    dist.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines))
        tot = torch.tensor([res["both_each_way"]], device="cuda")
    t = torch.tensor([res["h2d_only"], res["d2h_only"], res["both_each_way"]], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"sum over {world} ranks: h2d_only {t[0]:.1f}  d2h_only {t[1]:.1f}  both_each_way {t[2]:.1f} GB/s")
    dist.destroy_process_group()
else:
    print(line)
