"""Multi-GPU layout of the render path: views are independent, so they are sharded across ranks with NO collective
in the data path; the only exchange is a gather of per-view metrics (PSNR) or images at
the end.  One process per GPU, torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of view indices [0, num_views) over `world` ranks (sizes differ by at most 1)."""
    base, extra = divmod(num_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def interleave_views(num_views: int, rank: int, world: int) -> range:
    """Round-robin split: rank r takes views r, r + world, r + 2 world, ...  For views along a trajectory this gives every
    rank the same mix of the trajectory (neighbouring views cost about the same), where contiguous blocks give each rank one
    stretch of it: the ranks' workloads then differ by what their stretch happens to look at, and a synchronous job runs at
    the pace of the heaviest stretch."""
    return range(rank, num_views, world)


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


class SharedCloudUploader:
    """Host -> device upload of a cloud EVERY rank needs (views of one scene sharded over the ranks of a box).

    Naively each of the N ranks pushes the whole cloud over its own PCIe link (C2: 170 MB per rank, eight ranks sharing
    two NUMA nodes' memory controllers: measured e2e efficiency 0.84 at N = 8 in round 1).  Here rank r copies only rows
    [r P/N, (r+1) P/N) of every per-Gaussian array from (pinned) host memory into its slice of a device buffer and ONE
    all-gather per array (NCCL over NVLink / NVSwitch on GPUs, in place: the send slice lives inside the receive buffer)
    hands every rank the rest -- PCIe carries 1/N of the bytes, NVLink (900 GB/s per direction) the remainder.
    Arrays are given as {name: host tensor [P, ...]}; P is padded up to a multiple of N internally."""

    def __init__(self, host: dict, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.host = host
        self.P = next(iter(host.values())).shape[0]
        self.chunk = -(-self.P // self.world)
        self.dev = {k: torch.empty((self.chunk * self.world, *v.shape[1:]), dtype=v.dtype, device=device)
                    for k, v in host.items()}
        self.bytes_per_step = sum(min(self.chunk, max(0, self.P - self.rank * self.chunk)) * v[0].numel() * v.element_size()
                                  for v in host.values())

    def upload(self) -> dict:
        """Enqueues the copies and the all-gathers on the current stream; returns {name: device tensor [P, ...]}."""
        lo, hi = self.rank * self.chunk, min(self.P, (self.rank + 1) * self.chunk)
        out = {}
        for k, h in self.host.items():
            d = self.dev[k]
            if hi > lo:
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
            if self.world > 1:
                flat = d.view(-1)
                n = flat.numel() // self.world
                mine = flat[self.rank * n:(self.rank + 1) * n]
                if d.device.type != "cuda":
                    mine = mine.clone()   # gloo (CPU tests) does not take an input that aliases the output
                dist.all_gather_into_tensor(flat, mine, group=self.group)
            out[k] = d[: self.P]
        return out


class numa_local_allocation:
    """Context manager: while it is active the process runs on the CPUs NVML reports as local to the GPU, so that the
    pinned host buffers allocated inside it (first touch) land in that socket's memory -- on an 8-GPU box GPUs 0-3 and 4-7
    hang off different sockets, and a rank that stages its uploads in the other socket's memory pays the inter-socket hop on
    every copy.  The previous affinity is restored on exit: the process itself stays free to run wherever the host has an
    idle core (pinning the CPUs for good was measured to hurt on a shared host: a rank confined to a busy socket enqueues
    its launches late)."""

    def __init__(self, device_index: int):
        self.device_index, self.saved, self.cpus = device_index, None, []

    def __enter__(self):
        try:
            self.saved = os.sched_getaffinity(0)
        except (AttributeError, OSError):
            return self
        self.cpus = bind_to_gpu_numa_node(self.device_index)
        return self

    def __exit__(self, *exc):
        if self.saved is not None and self.cpus:
            try:
                os.sched_setaffinity(0, self.saved)
            except OSError:
                pass
        return False


def bind_to_gpu_numa_node(device_index: int) -> list:
    """Pins this process (and so the pinned host buffers it allocates afterwards: first touch) to the CPUs NVML reports as
    local to the GPU.  Returns the CPU list in effect ([] = unchanged).  See numa_local_allocation."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].strip().isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = local & set(os.sched_getaffinity(0))
        if allowed and allowed != set(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return sorted(allowed)
    except Exception:  # noqa: BLE001 -- no NVML, no permission, cpuset without local CPUs: leave the affinity alone
        pass
    return []
