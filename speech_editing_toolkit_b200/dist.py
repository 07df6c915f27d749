"""Batch-axis sharding of the sampling path over the GPUs of one box.

Utterances are independent, so rank r processes the contiguous slice [r*B/W, (r+1)*B/W) with replicated
weights and no collective inside the sampling loop; the only exchange is one all_gather of the generated
mels (and optionally wavs) at the end.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced slice of n items for `rank` (first n % world ranks get one extra item)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, rank: int, world: int):
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def all_gather_batch(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank results along dim 0 (ragged shards are padded to the largest shard, then trimmed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], 0)


def run_sharded(fn: Callable[..., torch.Tensor], batch_inputs: Sequence[torch.Tensor], group=None) -> torch.Tensor:
    """result = concat_r fn(*[x[shard_r] for x in batch_inputs]); every rank returns the full result."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = batch_inputs[0].shape[0]
    local = fn(*[shard(x, rank, world) for x in batch_inputs])
    return all_gather_batch(local, n, group)
