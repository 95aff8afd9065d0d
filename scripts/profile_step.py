"""Runs a few forward(+backward) steps of the bench workload (C2) with nothing else, for ncu captures:
  ncu --set full --clock-control none --import-source on -k regex:k_composite_fwd -s 2 -c 1 -o gpurun_out/x python scripts/profile_step.py
Also prints a pinned-host -> device copy bandwidth probe (context for bench.py's e2e number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200.cameras import make_view_batch  # noqa: E402
from pf3plat_b200.rasterizer import BatchSettings, rasterize_batch  # noqa: E402
from pf3plat_b200.synthetic import make_scene, make_target  # noqa: E402

P, V, HW = int(os.environ.get("GS_P", 500_000)), int(os.environ.get("GS_V", 8)), int(os.environ.get("GS_HW", 256))
steps = int(os.environ.get("GS_STEPS", 3))
bwd = os.environ.get("GS_BWD", "1") == "1"
dev = torch.device("cuda:0")
sc = make_scene(P, V, HW, HW, seed=0).to(dev)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
bs = BatchSettings(image_height=HW, image_width=HW, viewmatrix=vb.viewmatrix, projmatrix=vb.projmatrix, campos=vb.campos,
                   bg=sc.background, sh_degree=4, tanfov=vb.tanfov, tuning=int(os.environ.get("GS_TUNING", 0)))
means = sc.means[None].clone().requires_grad_(bwd)
opac = sc.opacities[None].clone().requires_grad_(bwd)
shs = sc.harmonics.permute(0, 2, 1).contiguous()[None].requires_grad_(bwd)
c = sc.covariances
cov6 = torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)[None].requires_grad_(bwd)
target = make_target(V, HW, HW).to(dev)
for _ in range(steps):
    color, radii = rasterize_batch(bs, means, opac, shs=shs, cov3D_precomp=cov6)
    if bwd:
        ((color - target) ** 2).mean().backward()
torch.cuda.synchronize()
if os.environ.get("GS_H2D", "0") == "1":
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"pinned H2D: {5 * h.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
    e0.record()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"pinned D2H: {5 * h.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
print("done")
