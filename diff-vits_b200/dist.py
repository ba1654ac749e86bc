"""Multi-GPU plumbing: independent utterances are batch-sharded across ranks (one process per GPU); the only
collective on the path is one all-gather of the output latents (the reference has no cross-sample reduction, so no
data-path exchange exists: SURVEY.md section 8e). Works with the NCCL backend on GPUs and with gloo on CPU (tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous shard of the batch for `rank`; the remainder is spread over the first ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_inputs(tensors: Tuple[torch.Tensor, ...], rank: int, world: int) -> Tuple[torch.Tensor, ...]:
    """Slice globally generated inputs (noise drawn once for the whole batch, then sharded, so the result is
    bit-identical to a single-GPU run)."""
    sl = shard_slice(tensors[0].shape[0], rank, world)
    return tuple(t[sl].contiguous() for t in tensors)


def gather_latents(latent: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather the per-rank output latents [B_r, C, T] into [global_batch, C, T] (ragged shards are padded)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return latent
    world = dist.get_world_size()
    per = (global_batch + world - 1) // world
    buf = latent
    if latent.shape[0] != per:
        buf = latent.new_zeros((per,) + tuple(latent.shape[1:]))
        buf[: latent.shape[0]] = latent
    out = latent.new_empty((world * per,) + tuple(latent.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous())
    if world * per == global_batch:
        return out
    parts = []
    for r in range(world):
        sl = shard_slice(global_batch, r, world)
        parts.append(out[r * per: r * per + (sl.stop - sl.start)])
    return torch.cat(parts, dim=0)
