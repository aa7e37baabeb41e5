"""Game sharding and statistics reduction for multi-GPU runs.

Leaf evaluations are independent and nothing shards mid-network, so the reference's multi-GPU mode
is "one model replica per GPU + per-job least-loaded routing" (reference
src/deepgo/native/cpp/Processor.cpp:17-60).  Here one PROCESS drives one GPU and owns a static
shard of the games / positions (game g -> rank g % world); the only collective is a reduction of a
few counters and of the step time (max over ranks) when a report is produced — NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List, Optional


def shard_items(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of the games / positions rank `rank` owns: {rank, rank + world, ...}."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world {world}')
    return list(range(rank, n_items, world))


@dataclass
class RankStats:
    positions: int = 0        # NN evaluations performed by this rank
    visits: int = 0           # search visits (self-play)
    batches: int = 0
    device_ms: float = 0.0    # CUDA-event time of the timed region on this rank
    wall_s: float = 0.0


def reduce_stats(stats: RankStats, dist=None, device=None) -> Dict[str, float]:
    """Whole-job aggregate: counters are summed, times are the MAX over ranks (the job is as slow as
    its slowest rank).  `dist` is torch.distributed (initialised) or None for a single process."""
    d = asdict(stats)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(d, world=1)
    import torch
    sums = torch.tensor([d['positions'], d['visits'], d['batches']], dtype=torch.float64, device=device)
    maxs = torch.tensor([d['device_ms'], d['wall_s']], dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    return dict(positions=int(sums[0].item()), visits=int(sums[1].item()), batches=int(sums[2].item()),
                device_ms=float(maxs[0].item()), wall_s=float(maxs[1].item()), world=dist.get_world_size())


def evals_per_sec(agg: Dict[str, float]) -> float:
    return agg['positions'] / (agg['device_ms'] / 1e3) if agg['device_ms'] > 0 else 0.0
