"""Per-stage times of the bench workload (C2 by default; GS_P / GS_V / GS_HW override), forward and backward:
the median over GS_STEPS steps of the CUDA-event brackets gs_set_profiling() puts around every stage, plus the
un-instrumented forward time (events around back-to-back calls)."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200.cameras import make_view_batch  # noqa: E402
from pf3plat_b200.rasterizer import BatchSettings, last_stats, rasterize_batch, set_profiling, stage_ms  # noqa: E402
from pf3plat_b200.synthetic import make_pixel_aligned_scene, make_scene, make_target  # noqa: E402

P, V, HW = int(os.environ.get("GS_P", 500_000)), int(os.environ.get("GS_V", 8)), int(os.environ.get("GS_HW", 256))
steps = int(os.environ.get("GS_STEPS", 20))
tuning = int(os.environ.get("GS_TUNING", 0))
dev = torch.device("cuda:0")
if os.environ.get("GS_SCENE") == "aligned":   # PF3plat-shaped: one Gaussian per pixel of two context views (P is ignored)
    sc = make_pixel_aligned_scene(HW, HW, V, seed=0).to(dev)
    P = sc.means.shape[0]
else:
    sc = make_scene(P, V, HW, HW, seed=0).to(dev)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
bs = BatchSettings(image_height=HW, image_width=HW, viewmatrix=vb.viewmatrix, projmatrix=vb.projmatrix, campos=vb.campos,
                   bg=sc.background, sh_degree=4, tanfov=vb.tanfov, tuning=tuning)
c = sc.covariances
cov6 = torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)[None]
shs = sc.harmonics.permute(0, 2, 1).contiguous()[None]
target = make_target(V, HW, HW).to(dev)


def step(grad):
    means = sc.means[None].clone().requires_grad_(grad)
    opac = sc.opacities[None].clone().requires_grad_(grad)
    s = shs.clone().requires_grad_(grad)
    cv = cov6.clone().requires_grad_(grad)
    color, _ = rasterize_batch(bs, means, opac, shs=s, cov3D_precomp=cv)
    if grad:
        ((color - target) ** 2).mean().backward()


for _ in range(3):
    step(True)
set_profiling(True, dev)
acc = {}
for _ in range(steps):
    step(True)
    torch.cuda.synchronize()
    for k, v in stage_ms(dev).items():
        acc.setdefault(k, []).append(v)
set_profiling(False, dev)
print({k: round(statistics.median(v), 4) for k, v in acc.items()})
with torch.no_grad():
    for _ in range(3):
        rasterize_batch(bs, sc.means[None], sc.opacities[None], shs=shs, cov3D_precomp=cov6)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        rasterize_batch(bs, sc.means[None], sc.opacities[None], shs=shs, cov3D_precomp=cov6)
    e1.record()
    torch.cuda.synchronize()
print(f"forward {e0.elapsed_time(e1) / steps:.4f} ms/step", last_stats(dev))
