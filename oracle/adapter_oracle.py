"""CPU restatement of the reference's Gaussian adapter, for the parity tests of the fused adapter kernels.

TEST INFRASTRUCTURE ONLY (imported by tests/ and the golden-vector script, never by the product path).

Follows, line by line, plain PyTorch code that IS in the reference tree:
  /root/reference/src/model/encoder/common/gaussian_adapter.py:48-111  (GaussianAdapter.forward, get_scale_multiplier)
  /root/reference/src/model/encoder/common/gaussians.py:8-45           (quaternion_to_matrix, build_covariance)
  /root/reference/src/geometry/projection.py:75-121                    (unproject, get_world_rays)
  /root/reference/src/misc/sh_rotation.py:10-36                        (rotate_sh, with the Wigner-D matrices as an input)
PINNED: tests/golden/make_adapter_golden.py runs the reference module itself (imported from /root/reference with stubs
for the absent jaxtyping / e3nn packages) and commits its outputs and autograd gradients as
tests/golden/adapter_*.npz; tests/test_adapter_cpu.py checks this restatement against them.  The one part the
reference delegates to e3nn -- the Wigner-D matrices -- enters both sides as data (pf3plat_b200/sh_rotation.py).

Everything is differentiable torch code (use float64 tensors for a tight reference); gradients come from autograd.
"""
from __future__ import annotations

from math import isqrt

import torch


def sh_mask(sh_degree: int) -> torch.Tensor:
    """gaussian_adapter.py:39-46."""
    m = torch.ones(((sh_degree + 1) ** 2,), dtype=torch.float32)
    for degree in range(1, sh_degree + 1):
        m[degree ** 2:(degree + 1) ** 2] = 0.1 * 0.25 ** degree
    return m


def quaternion_to_matrix(q: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """gaussians.py:8-31 (xyzw order)."""
    i, j, k, r = torch.unbind(q, dim=-1)
    two_s = 2 / ((q * q).sum(dim=-1) + eps)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(*q.shape[:-1], 3, 3)


def build_covariance(scale: torch.Tensor, rotation_xyzw: torch.Tensor) -> torch.Tensor:
    """gaussians.py:34-45."""
    s = scale.diag_embed()
    r = quaternion_to_matrix(rotation_xyzw)
    return r @ s @ s.transpose(-1, -2) @ r.transpose(-1, -2)


def get_scale_multiplier(intrinsics: torch.Tensor, pixel_size: torch.Tensor, multiplier: float = 0.1) -> torch.Tensor:
    """gaussian_adapter.py:100-111."""
    xy = multiplier * torch.einsum("...ij,j->...i", intrinsics[..., :2, :2].inverse(), pixel_size)
    return xy.sum(dim=-1)


def get_world_rays(coordinates, extrinsics, intrinsics):
    """projection.py:96-121 (with unproject, :75-93, inlined at z = 1)."""
    hom = torch.cat([coordinates, torch.ones_like(coordinates[..., :1])], dim=-1)
    directions = torch.einsum("...ij,...j->...i", intrinsics.inverse(), hom)
    directions = directions / directions.norm(dim=-1, keepdim=True)
    directions = torch.cat([directions, torch.zeros_like(directions[..., :1])], dim=-1)
    directions = torch.einsum("...ij,...j->...i", extrinsics, directions)[..., :-1]
    origins = extrinsics[..., :-1, -1].broadcast_to(directions.shape)
    return origins, directions


def adapter_forward(extrinsics, intrinsics, coordinates, depths, opacities, raw_gaussians, image_shape, sh_degree,
                    scale_min, scale_max, sh_rotation=None, eps: float = 1e-8) -> dict:
    """gaussian_adapter.py:48-98.  `sh_rotation`: block-diagonal (..., d_sh, d_sh) Wigner-D matrices broadcastable
    against the Gaussians' batch dims (what rotate_sh builds internally, sh_rotation.py:26-34), or None for identity."""
    d_sh = (sh_degree + 1) ** 2
    dt = raw_gaussians.dtype
    scales, rotations, sh = raw_gaussians.split((3, 4, 3 * d_sh), dim=-1)
    scales = scale_min + (scale_max - scale_min) * scales.sigmoid()
    h, w = image_shape
    pixel_size = 1 / torch.tensor((w, h), dtype=dt, device=raw_gaussians.device)
    multiplier = get_scale_multiplier(intrinsics, pixel_size)
    scales = scales * depths[..., None] * multiplier[..., None]
    rotations = rotations / (rotations.norm(dim=-1, keepdim=True) + eps)
    sh = sh.reshape(*sh.shape[:-1], 3, d_sh)
    sh = sh.broadcast_to((*opacities.shape, 3, d_sh)) * sh_mask(sh_degree).to(raw_gaussians)
    covariances = build_covariance(scales, rotations)
    c2w = extrinsics[..., :3, :3].detach()
    covariances = c2w @ covariances @ c2w.transpose(-1, -2)
    origins, directions = get_world_rays(coordinates, extrinsics, intrinsics)
    means = origins + directions * depths[..., None]
    if sh_rotation is not None:
        parts = []
        for degree in range(isqrt(d_sh)):
            lo, hi = degree ** 2, (degree + 1) ** 2
            parts.append(torch.einsum("...ij,...j->...i", sh_rotation[..., None, lo:hi, lo:hi].to(dt), sh[..., lo:hi]))
        sh = torch.cat(parts, dim=-1)
    return dict(means=means, covariances=covariances, harmonics=sh, opacities=opacities, scales=scales,
                rotations=rotations.broadcast_to((*scales.shape[:-1], 4)))
