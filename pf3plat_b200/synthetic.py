"""Deterministic synthetic Gaussian clouds and cameras.

The distributions resemble what PF3plat's adapter emits
(/root/reference/src/model/encoder/common/gaussian_adapter.py:60-98,
/root/reference/config/model/encoder/costvolume.yaml:13-16) and are insensitive
to the unknowns of SURVEY.md Appendix C (all depths >= 1.5, SH coefficients
16..24 zero).  Everything is generated on the CPU in fp32 from a seeded
torch.Generator, in the layout `render_cuda` takes
(/root/reference/src/model/decoder/cuda_splatting.py:47-60).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

FX = FY = 0.86  # normalised focal length => tanfov = 0.5/0.86
NEAR, FAR = 1.0, 100.0  # /root/reference/config/experiment/re10k.yaml:41-42


@dataclass
class Scene:
    extrinsics: torch.Tensor      # (V,4,4) camera-to-world
    intrinsics: torch.Tensor      # (V,3,3) normalised
    near: torch.Tensor            # (V,)
    far: torch.Tensor             # (V,)
    image_shape: tuple            # (h,w)
    background: torch.Tensor      # (V,3)
    means: torch.Tensor           # (P,3)
    covariances: torch.Tensor     # (P,3,3)
    harmonics: torch.Tensor       # (P,3,d_sh)
    opacities: torch.Tensor       # (P,)
    scales: torch.Tensor          # (P,3)   (covariances == R diag(scales^2) R^T)
    rotations: torch.Tensor       # (P,4)   quaternion (w,x,y,z), normalised

    def to(self, device):
        kw = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()}
        return Scene(**kw)


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    r, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(*q.shape[:-1], 3, 3)


def make_cameras(num_views: int, h: int, w: int, first_view: int = 0, total_views: int | None = None,
                 phase: float = 0.0, view_stride: int = 1):
    """View 0 = identity; view k = translated on a circle of radius 0.1 in the image plane
    (same construction as /root/reference/src/visualization/camera_trajectory/wobble.py:8-22).  The i-th camera returned
    is view first_view + i * view_stride of the total_views on the circle."""
    total = total_views or num_views
    ext = torch.eye(4, dtype=torch.float32).repeat(num_views, 1, 1)
    for i in range(num_views):
        k = first_view + i * view_stride
        if k > 0:
            ang = 2 * math.pi * k / total + phase
            ext[i, 0, 3] = 0.1 * math.cos(ang)
            ext[i, 1, 3] = 0.1 * math.sin(ang)
    K = torch.eye(3, dtype=torch.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = FX, FY, 0.5, 0.5
    intr = K.repeat(num_views, 1, 1)
    near = torch.full((num_views,), NEAR)
    far = torch.full((num_views,), FAR)
    bg = torch.zeros(num_views, 3)  # /root/reference/config/dataset/re10k.yaml:10
    return ext, intr, near, far, bg


def make_scene(num_gaussians: int, num_views: int, h: int, w: int, seed: int = 0, d_sh: int = 25,
               first_view: int = 0, total_views: int | None = None, view_stride: int = 1) -> Scene:
    g = torch.Generator().manual_seed(seed)
    P = num_gaussians

    def U(lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, generator=g)

    ndc = U(-1.05, 1.05, P, 2)
    depth = torch.exp(U(math.log(1.5), math.log(20.0), P))
    # unproject through view 0 (identity c2w): x_cam = ndc * tanfov * depth
    tanfov = 0.5 / FX
    means = torch.stack([ndc[:, 0] * tanfov * depth, ndc[:, 1] * tanfov * depth, depth], dim=-1)
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    mult = 0.1 * (1.0 / (FX * w) + 1.0 / (FY * h))
    scales = depth[:, None] * mult * torch.exp(U(math.log(0.5), math.log(15.0), P, 3))
    R = quat_to_rotmat(q)
    cov = R @ torch.diag_embed(scales * scales) @ R.transpose(-1, -2)
    cov = 0.5 * (cov + cov.transpose(-1, -2))
    opac = U(0.05, 1.0, P)
    mask = torch.ones(d_sh)
    for deg in range(1, int(math.isqrt(d_sh))):
        mask[deg * deg:(deg + 1) * (deg + 1)] = 0.1 * 0.25 ** deg
    if d_sh > 16:
        mask[16:] = 0.0
    sh = torch.randn(P, 3, d_sh, generator=g) * mask
    ext, intr, near, far, bg = make_cameras(num_views, h, w, first_view, total_views, view_stride=view_stride)
    return Scene(ext, intr, near, far, (h, w), bg, means.contiguous(), cov.contiguous(), sh.contiguous(),
                 opac.contiguous(), scales.contiguous(), q.contiguous())


def make_target(num_views: int, h: int, w: int, seed: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(num_views, 3, h, w, generator=g)


def make_pixel_aligned_scene(h: int, w: int, num_views: int, seed: int = 0, d_sh: int = 25, context_views: int = 2) -> Scene:
    """PF3plat-shaped cloud: one Gaussian per pixel of each context view (2*h*w Gaussians for the default two views,
    /root/reference/src/model/encoder/encoder_costvolume.py:556-573), unprojected along that pixel's ray at a smooth
    random depth, scales from the adapter's rule (gaussian_adapter.py:63-70).  Consecutive indices are neighbouring
    pixels, so -- unlike make_scene -- neighbouring threads of the kernels touch the same tiles."""
    g = torch.Generator().manual_seed(seed)
    tanfov = 0.5 / FX
    # target cameras never coincide with a context camera (first_view=1 skips the identity pose) nor are displaced
    # along a pixel axis only (phase): seen through such a camera the Gaussians project EXACTLY onto pixel rows or
    # columns, a measure-zero configuration in which footprints end exactly on tile boundaries and the discontinuous
    # decisions of the rasterizer become coin flips
    ext, intr, near, far, bg = make_cameras(num_views, h, w, first_view=1, total_views=num_views + 1, phase=0.4)
    ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
    means, scales_all = [], []
    for cv in range(context_views):
        # smooth depth map: a few random low-frequency cosines, 2..12 units deep
        ph = torch.rand(4, 3, generator=g) * 6.28
        fr = torch.rand(4, 2, generator=g) * 3 + 0.5
        depth = 5.0 + sum(1.2 * torch.cos(fr[k, 0] * 6.28 * xs + fr[k, 1] * 6.28 * ys + ph[k, 0]) for k in range(4))
        depth = depth.clamp_min(2.0) + 0.05 * torch.rand(h, w, generator=g)
        shift = 0.15 * cv                                   # second context camera displaced along x
        x = (2 * xs - 1) * tanfov * depth + shift
        y = (2 * ys - 1) * tanfov * depth
        means.append(torch.stack([x, y, depth], -1).reshape(-1, 3))
        mult = 0.1 * (1.0 / (FX * w) + 1.0 / (FY * h))
        sc = depth.reshape(-1, 1) * mult * (0.5 + 14.5 * torch.sigmoid(torch.randn(h * w, 3, generator=g)))
        scales_all.append(sc)
    means = torch.cat(means)
    scales = torch.cat(scales_all)
    P = means.shape[0]
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    R = quat_to_rotmat(q)
    cov = R @ torch.diag_embed(scales * scales) @ R.transpose(-1, -2)
    cov = 0.5 * (cov + cov.transpose(-1, -2))
    opac = torch.sigmoid(torch.randn(P, generator=g))
    mask = torch.ones(d_sh)
    for deg in range(1, int(math.isqrt(d_sh))):
        mask[deg * deg:(deg + 1) * (deg + 1)] = 0.1 * 0.25 ** deg
    if d_sh > 16:
        mask[16:] = 0.0
    sh = torch.randn(P, 3, d_sh, generator=g) * mask
    return Scene(ext, intr, near, far, (h, w), bg, means.contiguous(), cov.contiguous(), sh.contiguous(),
                 opac.contiguous(), scales.contiguous(), q.contiguous())
