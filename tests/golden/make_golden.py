"""Generates tests/golden/*.npz from the CPU oracle (fp32 build) on the seeded synthetic scenes of SURVEY.md
section 8(d).  The reference ships no golden vectors for the rasterizer and its CUDA
extension is absent, so these are ORACLE outputs, committed to (1) freeze the oracle against regressions and
(2) give the GPU tests a reference that needs no CPU recomputation.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.gs_oracle import OracleRender, OracleSettings  # noqa: E402
from pf3plat_b200.synthetic import make_scene, make_target  # noqa: E402
from tests.util import view_args  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scene_case(name, P, views, h, w, seed, use_sh=True):
    sc = make_scene(P, views, h, w, seed=seed)
    out = {"P": P, "views": views, "h": h, "w": w, "seed": seed, "use_sh": use_sh}
    target = make_target(views, h, w).numpy()
    for v in range(views):
        st, kw = view_args(sc, v, use_sh=use_sh)
        r = OracleRender(st, **kw, with_depth=True)
        dL = (2 * (r.color - target[v]) / target.size).astype(np.float32)
        g = r.backward(dL)
        out[f"color{v}"] = r.color
        out[f"depth{v}"] = r.depth
        out[f"radii{v}"] = r.radii
        out[f"fragile{v}"] = np.packbits(r.px_fragile)
        out[f"geom_fragile{v}"] = np.packbits(r.geom_fragile > 0)
        out[f"g_means{v}"] = g["means3D"]
        out[f"g_opac{v}"] = g["opacities"]
        out[f"g_cov{v}"] = g["cov3D_precomp"]
        out[f"num_rendered{v}"] = r.num_rendered
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written")


def single_gaussian_spin():
    """The scene of /root/reference/src/scripts/test_splatter.py:30-65 in miniature: ONE Gaussian at the origin,
    identity covariance, degree-2 SH coefficients set to 10 on the red channel, seen from 4 poses on a radius-10
    circle about the y axis (near 0.1 / far 20, hence the x10 scale-invariant rescale)."""
    h = w = 64
    near, far = 0.1, 20.0
    scale = 1.0 / near
    K_tan = 0.5 / 0.5  # intrinsics [[.5,0,.5],[0,.5,.5]] -> tan(fov/2) = 1
    sh = np.zeros((1, 25, 3), np.float32)
    sh[0, 4:9, 0] = 10.0
    imgs = []
    for k in range(4):
        a = 2 * np.pi * k / 4
        c2w = np.eye(4)
        c2w[:3, :3] = [[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]]
        c2w[:3, 3] = c2w[:3, :3] @ np.array([0, 0, -10.0])
        c2w[:3, 3] *= scale
        view = np.linalg.inv(c2w).T
        n_, f_ = near * scale, far * scale
        proj = np.zeros((4, 4))
        proj[0, 0] = 1 / K_tan
        proj[1, 1] = 1 / K_tan
        proj[3, 2] = 1
        proj[2, 2] = f_ / (f_ - n_)
        proj[2, 3] = -(f_ * n_) / (f_ - n_)
        st = OracleSettings(image_height=h, image_width=w, tanfovx=K_tan, tanfovy=K_tan, bg=np.zeros(3),
                            scale_modifier=1.0, viewmatrix=view, projmatrix=view @ proj.T, sh_degree=4,
                            campos=c2w[:3, 3])
        r = OracleRender(st, means3D=np.zeros((1, 3)), opacities=np.ones(1), shs=sh,
                         cov3D_precomp=np.array([[1, 0, 0, 1, 0, 1.0]]) * scale * scale)
        imgs.append(r.color)
    np.savez_compressed(os.path.join(HERE, "single_gaussian_spin.npz"), color=np.stack(imgs))
    print("single_gaussian_spin written")


if __name__ == "__main__":
    scene_case("scene_sh_2k", 2000, 2, 48, 64, seed=11)
    scene_case("scene_rgb_1k", 1000, 1, 40, 24, seed=12, use_sh=False)
    single_gaussian_spin()
