"""Per-step wall time of forward+backward in the usual training-loop pattern (the previous step's output stays bound until
the new forward has returned), default at the C4 size (2M Gaussians, 32 views, 512x512; GS_P / GS_V / GS_HW), with the
library's counters."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pf3plat_b200 import rasterizer
from pf3plat_b200.cameras import make_view_batch
from pf3plat_b200.rasterizer import BatchSettings, rasterize_batch
from pf3plat_b200.synthetic import make_scene, make_target
dev = torch.device("cuda:0")
P, V, HW = int(os.environ.get("GS_P", 2_000_000)), int(os.environ.get("GS_V", 32)), int(os.environ.get("GS_HW", 512))
sc = make_scene(P, V, HW, HW, seed=0)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far, scale_invariant=True)
c = sc.covariances
d = {"means3D": sc.means.reshape(1, P, 3), "opacities": sc.opacities.reshape(1, P), "shs": sc.harmonics.permute(0, 2, 1).contiguous().reshape(1, P, 25, 3),
     "cov3D_precomp": torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1).reshape(1, P, 6)}
leaves = {k: v.contiguous().float().to(dev).requires_grad_(True) for k, v in d.items()}
bs = BatchSettings(image_height=HW, image_width=HW, viewmatrix=vb.viewmatrix.to(dev), projmatrix=vb.projmatrix.to(dev), campos=vb.campos.to(dev),
                   bg=sc.background.to(dev), sh_degree=4, tanfov=vb.tanfov.to(dev))
target = make_target(V, HW, HW, seed=5).to(dev)
for it in range(int(os.environ.get("GS_STEPS", 14))):
    if it == 4 and os.environ.get("GS_NOGRAD_BETWEEN"):
        with torch.no_grad():
            for _ in range(3):
                rasterize_batch(bs, leaves["means3D"], leaves["opacities"], shs=leaves["shs"], cov3D_precomp=leaves["cov3D_precomp"])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in leaves.values():
        t.grad = None
    col, _ = rasterize_batch(bs, leaves["means3D"], leaves["opacities"], shs=leaves["shs"], cov3D_precomp=leaves["cov3D_precomp"])
    t1 = time.perf_counter()
    ((col - target) ** 2).mean().backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    s = rasterizer.last_stats(dev)
    print(f"step {it}: fwd enqueue {1e3*(t1-t0):7.2f} ms, total {1e3*(t2-t0):7.2f} ms  spec {s['speculative']} redos {s['overflow_redos']} saved {s['saved_bytes']>>20} MB scratch {s['scratch_bytes']>>20} MB  pool reserved {s['pool_reserved_bytes']>>20} used {s['pool_used_bytes']>>20} MB")
