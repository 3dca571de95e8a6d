"""
Multi-GPU execution of batched rollouts: one process per GPU (`torchrun`), the batch axis is
sharded across ranks and every rank runs the same fused kernels on its slice.  Trajectories
never interact (the reference batches with `jax.vmap` only, exponax/_base_stepper.py:260-262), so
the data path needs NO collective; `torch.distributed` is used only to optionally gather results
and to agree on timings (max over ranks).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of the batch axis owned by `rank`: the first
    `batch % world_size` ranks get one extra trajectory."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} out of range for world size {world_size}")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_shard(u0):
    """This rank's slice of a replicated batched initial condition (leading axis = batch)."""
    rank, ws = world()
    lo, hi = shard_bounds(len(u0), ws, rank)
    return u0[lo:hi]


def sharded_apply(fn, u0, *, gather: bool = False):
    """Run `fn` (e.g. `ex.vmap(ex.rollout(stepper, T))`) on this rank's shard of `u0`.

    gather=False: returns the local result (the usual case: results stay where they were made).
    gather=True : all ranks receive the full result, concatenated along the batch axis in rank
                  order (the only collective; not part of the timed data path)."""
    rank, ws = world()
    lo, hi = shard_bounds(len(u0), ws, rank)
    local = fn(u0[lo:hi]) if hi > lo else None
    if not gather or ws == 1:
        return local
    is_np = isinstance(local, np.ndarray) or (local is None and isinstance(u0, np.ndarray))
    sizes = [shard_bounds(len(u0), ws, r) for r in range(ws)]
    # shapes can differ by one along the batch axis: exchange them first
    shape = [None] * ws
    dist.all_gather_object(shape, None if local is None else tuple(local.shape))
    ref_shape = next(s for s in shape if s is not None)
    backend = dist.get_backend()
    dev = "cuda" if backend == "nccl" else "cpu"
    lt = torch.as_tensor(local) if local is not None else None
    dtype = lt.dtype if lt is not None else torch.float32
    parts = []
    for r, (a, b) in enumerate(sizes):
        shp = (b - a,) + tuple(ref_shape[1:])
        buf = torch.empty(shp, dtype=dtype, device=dev)
        if r == rank and lt is not None:
            buf.copy_(lt)
        if b > a:
            dist.broadcast(buf, src=r)
        parts.append(buf)
    full = torch.cat(parts, dim=0)
    return full.cpu().numpy() if is_np else full


def max_over_ranks(value: float) -> float:
    """Device-agnostic max reduction used for multi-GPU timings."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
