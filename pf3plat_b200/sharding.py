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


def allreduce_scene_gradients(grads: list) -> None:
    """The one exchange step of the training case where ONE scene's views are split across ranks (SURVEY.md section
    8(e)): every rank holds the full Gaussian set, renders its own views forward+backward, and the per-Gaussian
    gradient blocks (means 3 + covariance 6 + SH 75 + opacity 1 floats per Gaussian) are summed over ranks, in place.
    The tensors are flattened into one bucket so that a single all-reduce (NCCL over NVLink on GPUs) moves them.
    PF3plat's own DDP shards by scene, where this is not needed (replicas only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    live = [g for g in grads if g is not None]
    if not live:
        return
    flat = torch.cat([g.reshape(-1) for g in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in live:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
