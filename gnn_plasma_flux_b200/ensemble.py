"""Ensemble sharding across GPUs: independent initial conditions, one process per GPU,
no data-path collective.  The reference loops ICs sequentially
(scripts/evaluation/evaluate_multi_ic.py:124-126); ICs never interact, so each rank
advances a contiguous block and only results/metrics are gathered at the end."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_ics: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of ICs owned by `rank`; blocks differ by at most one IC."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_ics, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_states(local: torch.Tensor, n_ics: int) -> torch.Tensor:
    """All ranks' final states [n_ics,3,nx] on every rank (result collection only, off the timed path).
    Uses the default process group; blocks may be ragged by one IC."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_ics, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], dim=0)
