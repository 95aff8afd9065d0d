"""Per-call host overhead of the drop-in operator at PF3plat's size (131072 Gaussians, 1 view, 256x256)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from pf3plat_b200.cameras import make_view_batch  # noqa: E402
from pf3plat_b200.synthetic import make_scene  # noqa: E402

dev = torch.device("cuda:0")
P = 131072
sc = make_scene(P, 1, 256, 256, seed=1).to(dev)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
row, col = torch.triu_indices(3, 3)
shs = sc.harmonics.permute(0, 2, 1).contiguous()
cov6 = sc.covariances[:, row, col]
opac = sc.opacities[:, None]
tx, ty = float(vb.tanfov[0, 0]), float(vb.tanfov[0, 1])


def call(grad=False):
    st = GaussianRasterizationSettings(image_height=256, image_width=256, tanfovx=tx, tanfovy=ty, bg=sc.background[0],
                                       scale_modifier=1.0, viewmatrix=vb.viewmatrix[0], projmatrix=vb.projmatrix[0],
                                       sh_degree=4, campos=vb.campos[0], prefiltered=False, debug=False)
    m = sc.means.requires_grad_(grad)
    img, radii = GaussianRasterizer(st)(means3D=m, means2D=torch.zeros_like(sc.means, requires_grad=grad), shs=shs,
                                        opacities=opac, cov3D_precomp=cov6)
    if grad:
        img.sum().backward()
    return img


for grad in (False, True):
    for _ in range(5):
        call(grad)
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        call(grad)
    torch.cuda.synchronize()
    print(f"drop-in op, grad={grad}: {1e6 * (time.perf_counter() - t0) / n:.0f} us per call")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    call(False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
