"""Probe: torch symmetric memory (NVLink peer mappings) between the ranks of one box."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty((world, 1024), dtype=torch.float64, device=torch.device("cuda", local))
t.fill_(-1.0)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0), flush=True)
hdl.barrier()
for peer in range(world):
    buf = hdl.get_buffer(peer, (world, 1024), torch.float64)
    buf[rank].fill_(float(rank))           # write my row into every peer's buffer
hdl.barrier()
torch.cuda.synchronize()
ok = all(float(t[r].min()) == float(r) == float(t[r].max()) for r in range(world))
print(rank, "peer writes visible:", ok, flush=True)
dist.destroy_process_group()
