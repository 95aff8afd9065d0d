"""Multi-GPU layout of the render path: views are independent, so they are sharded across ranks with NO collective
in the data path; the only exchange is a gather of per-view metrics (PSNR) or images at
the end.  One process per GPU, torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of view indices [0, num_views) over `world` ranks (sizes differ by at most 1)."""
    base, extra = divmod(num_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_metric(local: torch.Tensor) -> torch.Tensor:
    """All-gathers a 1-D per-view metric from every rank, concatenated in rank (= view) order.  Ragged shards are
    supported by padding to the longest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.numel()], device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)])
