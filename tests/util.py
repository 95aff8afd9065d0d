"""Shared helpers for the parity tests: turn a synthetic Scene into the exact per-view arguments the
reference call site hands to the rasterizer (/root/reference/src/model/decoder/cuda_splatting.py:64-124)."""
from __future__ import annotations

import math

import numpy as np
import torch

from oracle.gs_oracle import OracleRender, OracleSettings
from pf3plat_b200.cameras import make_view_batch


def view_args(scene, v: int, use_sh: bool = True):
    """Returns (OracleSettings, kwargs) for view v, following render_cuda line by line."""
    vb = make_view_batch(scene.extrinsics, scene.intrinsics, scene.near, scene.far, scale_invariant=True)
    scale = vb.scale[v]
    h, w = scene.image_shape
    means = scene.means * scale
    cov = scene.covariances * scale * scale
    row, col = torch.triu_indices(3, 3)
    cov6 = cov[:, row, col]
    d_sh = scene.harmonics.shape[-1]
    shs = scene.harmonics.permute(0, 2, 1).contiguous()  # (P, d_sh, 3)
    st = OracleSettings(
        image_height=h, image_width=w, tanfovx=float(vb.tanfov[v, 0]), tanfovy=float(vb.tanfov[v, 1]),
        bg=scene.background[v].numpy(), scale_modifier=1.0, viewmatrix=vb.viewmatrix[v].numpy(),
        projmatrix=vb.projmatrix[v].numpy(), sh_degree=math.isqrt(d_sh) - 1, campos=vb.campos[v].numpy())
    kw = dict(means3D=means.numpy(), opacities=scene.opacities.numpy(), cov3D_precomp=cov6.numpy())
    if use_sh:
        kw["shs"] = shs.numpy()
    else:
        kw["colors_precomp"] = shs[:, 0, :].contiguous().numpy()
    return st, kw


def oracle_view(scene, v, use_sh=True, dtype=np.float32, **extra):
    st, kw = view_args(scene, v, use_sh)
    return OracleRender(st, dtype=dtype, **kw, **extra)


def simple_settings(h=64, w=64, tanfov=0.5, bg=(0.0, 0.0, 0.0), near=1.0, far=100.0, sh_degree=0):
    """Identity camera looking down +z, built the way get_projection_matrix does (cuda_splatting.py:17-44)."""
    proj = np.zeros((4, 4), np.float64)
    proj[0, 0] = 1.0 / tanfov
    proj[1, 1] = 1.0 / tanfov
    proj[3, 2] = 1.0
    proj[2, 2] = far / (far - near)
    proj[2, 3] = -(far * near) / (far - near)
    view = np.eye(4)
    full = view.T @ proj.T
    return OracleSettings(image_height=h, image_width=w, tanfovx=tanfov, tanfovy=tanfov, bg=np.array(bg, np.float64),
                          scale_modifier=1.0, viewmatrix=view.T.copy(), projmatrix=full, sh_degree=sh_degree,
                          campos=np.zeros(3))
