"""PF3plat-shaped training step over this repo's components (reduced):

    context images -> [stand-in encoder] -> raw Gaussians, depths, densities
                   -> pf3plat_b200.adapter.GaussianAdapter        (fused kernels, gaussian_adapter.py:48-98)
                   -> pf3plat_b200.render.decoder_forward         (one batched rasterizer call, colour + depth)
                   -> MSE loss (loss/loss_mse.py) -> backward -> DDP gradient all-reduce -> SGD step
                   -> PSNR / SSIM of the rendered target views    (pf3plat_b200.metrics)

The reference's own encoder (EncoderCostVolume: UniDepth, LightGlue, cost volume; /root/reference/src/model/encoder/
encoder_costvolume.py) cannot be instantiated here -- its third-party networks and weights are absent and north_star keeps
it as stock PyTorch -- so a small convolutional stand-in with the same OUTPUT contract (82 raw channels + depth + density
per pixel of each of two context views, encoder_costvolume.py:529-573) feeds the path.  Everything behind the encoder is
the real thing.  Sharding is PF3plat's: DDP by scene (main.py:104-116), one process per GPU, NCCL; the rasterizer needs
no collective ("replicas only").

  python scripts/c5_chain.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/c5_chain.py
Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200.adapter import GaussianAdapter, GaussianAdapterCfg  # noqa: E402
from pf3plat_b200.metrics import compute_psnr, compute_ssim  # noqa: E402
from pf3plat_b200.render import decoder_forward  # noqa: E402
from pf3plat_b200.synthetic import make_cameras  # noqa: E402

H = W = int(os.environ.get("C5_HW", 256))
SCENES_PER_RANK = int(os.environ.get("C5_SCENES", 1))   # re10k.yaml batch size per GPU
CONTEXT, TARGET = 2, int(os.environ.get("C5_TARGETS", 4))
STEPS, WARMUP = int(os.environ.get("C5_STEPS", 10)), 3
SH_DEGREE = 4


class StandInEncoder(nn.Module):
    """Same output contract as the reference encoder's Gaussian head: per context pixel 2 offset + 82 raw channels, a
    depth in [near, far] and a density in (0, 1)."""

    def __init__(self, d_in):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.GELU(), nn.Conv2d(32, 64, 3, padding=1), nn.GELU(),
                                 nn.Conv2d(64, 2 + d_in + 2, 1))

    def forward(self, images, near, far):
        b, v = images.shape[:2]
        out = self.net(images.flatten(0, 1)).reshape(b, v, -1, H * W).transpose(-1, -2)   # (b, v, hw, c)
        offset, raw, dd = out[..., :2], out[..., 2:-2], out[..., -2:]
        depth = near[..., None] + (far - near)[..., None] * torch.sigmoid(dd[..., 0]) * 0.1 + 1.0
        return offset, raw, depth, torch.sigmoid(dd[..., 1])


def main():
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)   # identical initial weights on every rank
    adapter = GaussianAdapter(GaussianAdapterCfg(0.5, 15.0, SH_DEGREE)).to(dev)
    enc = StandInEncoder(adapter.d_in).to(dev)
    model = nn.parallel.DistributedDataParallel(enc, device_ids=[local]) if world > 1 else enc
    opt = torch.optim.SGD(enc.parameters(), lr=1e-4)

    b, g = SCENES_PER_RANK, torch.Generator().manual_seed(100 + rank)    # every rank: its own scenes
    ext, intr, near, far, _ = make_cameras(CONTEXT + TARGET, H, W, first_view=1, total_views=CONTEXT + TARGET + 1, phase=0.4)
    rep = lambda t: t[None].repeat(b, *([1] * t.dim())).to(dev)
    ext, intr, near, far = rep(ext), rep(intr), rep(near), rep(far)
    images = torch.rand(b, CONTEXT + TARGET, 3, H, W, generator=g).to(dev)
    ys, xs = torch.meshgrid((torch.arange(H) + 0.5) / H, (torch.arange(W) + 0.5) / W, indexing="ij")
    xy = torch.stack([xs, ys], -1).reshape(1, 1, H * W, 2).to(dev)
    pixel = torch.tensor([1.0 / W, 1.0 / H], device=dev)
    bg = torch.zeros(3, device=dev)
    cx, tg = slice(0, CONTEXT), slice(CONTEXT, CONTEXT + TARGET)

    def step():
        offset, raw, depth, density = model(images[:, cx], near[:, cx], far[:, cx])
        coords = xy + (torch.sigmoid(offset) - 0.5) * pixel                                   # encoder_costvolume.py:514-516
        gs = adapter(ext[:, cx, None], intr[:, cx, None], coords, depth, density, raw, (H, W))  # (b, v, hw, ...)
        flat = lambda t, n: t.reshape(b, -1, *t.shape[-n:]) if n else t.reshape(b, -1)          # "b (v r) ..."
        color, dep = decoder_forward(flat(gs.means, 1), flat(gs.covariances, 2), flat(gs.harmonics, 2), flat(gs.opacities, 0),
                                     ext[:, tg], intr[:, tg], near[:, tg], far[:, tg], (H, W), bg, depth_mode="depth")
        loss = ((color - images[:, tg]) ** 2).mean()                                            # loss_mse.py
        opt.zero_grad(set_to_none=True)
        loss.backward()                                                                         # DDP all-reduces here
        opt.step()
        return loss.detach(), color.detach()

    for _ in range(WARMUP):
        loss, color = step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(STEPS):
        loss, color = step()
        losses.append(loss)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev)
    psnr = compute_psnr(images[:, tg].flatten(0, 1), color.flatten(0, 1)).mean()
    ssim = compute_ssim(images[:, tg].flatten(0, 1), color.flatten(0, 1)).mean()
    stats = torch.stack([ms[0], psnr, ssim, torch.stack(losses)[-1]])
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        stats /= world
    if rank == 0:
        finite = all(torch.isfinite(p.grad).all().item() for p in enc.parameters())
        print(json.dumps({"workload": f"C5 chain: {b} scene(s)/GPU x ({CONTEXT} context + {TARGET} target views) x {H}x{W}, "
                                      f"{CONTEXT * H * W} Gaussians/scene, stand-in encoder", "n_gpus": world,
                          "ms_per_step": float(ms[0]), "scenes_per_sec": b * world / (float(ms[0]) * 1e-3),
                          "loss_first": float(losses[0]), "loss_last": float(stats[3]), "psnr_db": float(stats[1]),
                          "ssim": float(stats[2]), "gradients_finite": finite, "parallelism": "DDP by scene (NCCL), no collective in the rasterizer"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
