"""Batch sharding of QKANLayer.forward over the GPUs of one box.

Samples are independent and the weights are shared read-only, so the path shards with no
data-path collective: rank r of g takes the contiguous slice [r*B/g, (r+1)*B/g) (remainder
spread over the first ranks), every rank runs the same kernel, and the only communication
is one gather of the [B/g, K] outputs (NCCL all_gather over NVLink/NVSwitch; gloo on CPU
for the host-logic tests).  The gathered result is bitwise equal to the single-GPU result.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of rank `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B: int, world: int) -> List[int]:
    return [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]


def gather_outputs(local: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All ranks receive the full [B, K] output.  Uneven shards are padded to the largest one
    so a single all_gather_into_tensor moves everything."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    mx = max(sizes) if sizes else 0
    K = local.shape[1]
    if local.is_cuda:
        torch.cuda.nvtx.range_push("qkan gather_outputs (NCCL all_gather)")
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0], K), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    full = torch.empty((world * mx, K), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    else:  # gloo: list form
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
        full = torch.cat(parts, dim=0)
    if local.is_cuda:
        torch.cuda.nvtx.range_pop()
    if all(s == mx for s in sizes):
        return full
    return torch.cat([full[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)


class FusedGatherQKANLayer:
    """Multi-GPU forward whose output gather is fused into the kernel: every rank's kernel stores its
    rows straight into the result buffer of every NVLink peer (torch symmetric memory gives the peer
    mappings), so there is no separate collective and the transfer overlaps the arithmetic.  After
    ``forward`` (which ends with a symmetric-memory barrier) every rank holds the full ``[B, K]``.

    One process per GPU, NCCL process group initialised, CUDA ``QKANLayer`` with ``prep="analytic"``.

    The result lives in symmetric memory that the peers write into.  Two buffers per shape alternate between
    calls, and every call ends with a barrier (``barrier=True``), so a rank that is already in call i+1 stores into
    the other buffer while a slower peer may still be reading the result of call i; the tensor returned by call i is
    overwritten by call i+2.  With ``barrier=False`` nothing orders the ranks: the caller must synchronise them
    (e.g. ``dist.barrier()``) before reading the result AND before the next-but-one call."""

    def __init__(self, layer, group=None, multicast: bool = True):
        self.layer = layer
        self.group = group if group is not None else dist.group.WORLD
        self.multicast = multicast          # use the NVSwitch multicast mapping when the fabric offers one
        self._bufs = {}
        self._calls = {}

    def _buffer(self, B: int, K: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        key = (B, K)
        if key not in self._bufs:
            pair = []
            for _ in range(2):
                t = symm_mem.empty((B, K), dtype=torch.float64, device=device)
                pair.append((t, symm_mem.rendezvous(t, self.group)))
            self._bufs[key] = pair
            self._calls[key] = 0
        self._calls[key] += 1
        return self._bufs[key][self._calls[key] & 1]

    def forward(self, x_local: torch.Tensor, weights, B_total: int, barrier: bool = True) -> torch.Tensor:
        """x_local: this rank's contiguous slice (shard_bounds) as a CUDA tensor [B_local, N]."""
        import ctypes
        from . import _binding as _b
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_bounds(B_total, world, rank)
        if x_local.shape[0] != hi - lo:
            raise ValueError(f"rank {rank} expects {hi - lo} rows, got {x_local.shape[0]}")
        if world > 8:
            raise ValueError("the fused gather addresses at most 8 peers (one NVSwitch box)")
        eng = self.layer._engine
        self.layer._set_weights(weights)
        out, hdl = self._buffer(B_total, self.layer.K, x_local.device)
        xd = x_local.to(torch.float64).contiguous()
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if self.multicast else 0
        self.last_path = "multicast" if mc else "peer-stores"
        if mc:
            # one multimem.st per result, replicated to every rank (this one included) by the NVSwitch
            _b.check(_b.lib().qkan_layer_forward_multicast(eng.handle(), xd.data_ptr(), xd.shape[0], ctypes.c_void_p(mc), lo,
                                                           eng._stream_ptr()))
        else:
            # own buffer first, then the peers
            order = [rank] + [r for r in range(world) if r != rank]
            ptrs = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(hdl.buffer_ptrs[r])) for r in order])
            _b.check(_b.lib().qkan_layer_forward_peers(eng.handle(), xd.data_ptr(), xd.shape[0], ptrs, world, lo,
                                                       eng._stream_ptr()))
        if barrier:
            torch.cuda.nvtx.range_push("qkan fused gather: symmetric-memory barrier")
            hdl.barrier()          # every rank's kernel has finished: all rows of all ranks are in place
            torch.cuda.nvtx.range_pop()
        return out


class ShardedQKANLayer:
    """One process per GPU; `forward` takes the FULL batch description and computes this
    rank's slice.  ``compute`` is the per-rank callable (the CUDA QKANLayer on a GPU box;
    injectable so the sharding logic can be tested on CPU with gloo)."""

    def __init__(self, layer=None, compute=None, group=None):
        if (layer is None) == (compute is None):
            raise ValueError("give exactly one of `layer` or `compute`")
        self.layer = layer
        self.compute = compute if compute is not None else (lambda x, w: layer.forward(x, w))
        self.group = group

    def forward_local(self, x_full: torch.Tensor, weights) -> Tuple[torch.Tensor, Tuple[int, int]]:
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_bounds(x_full.shape[0], world, rank)
        return self.compute(x_full[lo:hi], weights), (lo, hi)

    def forward(self, x_full: torch.Tensor, weights, gather: bool = True):
        y, _ = self.forward_local(x_full, weights)
        if not gather:
            return y
        return gather_outputs(y, x_full.shape[0], self.group)
