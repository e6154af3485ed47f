"""Multi-GPU host path: the batch of samples shards across ranks (one process per GPU,
full weight replica each, both CFG halves of an image on the same rank), there is no collective
inside the DDIM loop, and ONE all-gather collects the decoded images.  torch.distributed is
plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests); the sampling path never imports
torch."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous block partition of `total` images: rank r gets [lo, hi).  The remainder goes to
    the first ranks, so any total works (ragged shards are padded in allgather_images)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(global_array, rank: int, world: int, axis: int = 0):
    """Slice of a globally seeded tensor (x_T [B,h,w,4], or noise [S,B,h,w,4] with axis=1): every
    rank draws from the same global tensor, so results do not depend on the GPU count."""
    lo, hi = shard_range(global_array.shape[axis], rank, world)
    idx = [slice(None)] * global_array.ndim
    idx[axis] = slice(lo, hi)
    return np.ascontiguousarray(global_array[tuple(idx)])


def shard_token_ids(token_ids, rank: int, world: int):
    """get_token_ids layout is B uncond rows then B cond rows (run_ldm_sampler.py:42-45); a rank
    needs the uncond AND cond rows of its own images."""
    b = token_ids.shape[0] // 2
    lo, hi = shard_range(b, rank, world)
    return np.concatenate([token_ids[lo:hi], token_ids[b + lo:b + hi]], axis=0)


def allgather_images(local, total: int, group=None):
    """All ranks end up with the [total, H, W, 3] image tensor in global sample order.  `local` is
    a torch tensor (CUDA for NCCL, CPU for gloo) holding this rank's shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < maxn:  # ragged last shards: pad to a common size for the collective
        pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    out = torch.empty((world * maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(hi - lo == maxn for lo, hi in sizes):
        return out
    parts = [out[r * maxn: r * maxn + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
