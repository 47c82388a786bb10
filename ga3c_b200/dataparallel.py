"""Data-parallel plumbing for the A3C train step.

Every loss term of the reference graph is a `reduce_sum` over the batch (NetworkVP_discrate.py:61,
:83-85), so the gradient of the concatenated batch is the plain SUM of the per-rank gradients: one
allreduce(SUM, no division) over the flat gradient arena, then the identical RMSProp update on every
rank keeps the replicas bit-identical.  The arena is reduced in two segments so the big one overlaps
the conv backward:

    [0, split)            conv11/*, conv12/*, dense1/b, logits_*   -- final only after the conv backward
    [split, arena_floats) dense1/w (98.8 % of the parameters)      -- final after ga3c_fb_head

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous, near-equal split of a global batch: rows [lo, hi) belong to `rank`."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradientAllReduce:
    """SUM-allreduce of the gradient arena in two segments; `big` can run on a side stream."""

    def __init__(self, arena: torch.Tensor, split: int, group=None):
        if not (0 <= split <= arena.numel()):
            raise ValueError("split outside the arena")
        self.small = arena[:split]
        self.big = arena[split:]
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self._comm = None
        self._ev_ready = self._ev_done = None
        if self.enabled and arena.is_cuda:
            with torch.cuda.device(arena.device):
                self._comm = torch.cuda.Stream(device=arena.device)
                self._ev_ready = torch.cuda.Event()
                self._ev_done = torch.cuda.Event()

    def start_big(self, stream=None):
        """Call when the dense1/w gradient is final on `stream`: launches its allreduce on the side stream."""
        if not self.enabled:
            return
        if self._comm is None:                       # CPU tensors (gloo): synchronous
            dist.all_reduce(self.big, op=dist.ReduceOp.SUM, group=self.group)
            return
        self._ev_ready.record(stream)
        self._comm.wait_event(self._ev_ready)
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self.big, op=dist.ReduceOp.SUM, group=self.group)
            self._ev_done.record(self._comm)

    def finish(self, stream=None):
        """Call when every gradient is final on `stream`: reduces the small segment and joins the big one."""
        if not self.enabled:
            return
        if self._comm is None:
            dist.all_reduce(self.small, op=dist.ReduceOp.SUM, group=self.group)
            return
        with torch.cuda.stream(stream):
            dist.all_reduce(self.small, op=dist.ReduceOp.SUM, group=self.group)
        stream.wait_event(self._ev_done)


def exchange_ipc_handles(mine: bytes, device: torch.device, group=None) -> bytes:
    """all_gather of one fixed-size opaque handle per rank (CUDA IPC handle of the rank's state slab);
    returns the handles concatenated in rank order."""
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
