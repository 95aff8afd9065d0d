"""Camera glue of the render boundary, batched over views.

Restates what `render_cuda` does before its per-view loop
(/root/reference/src/model/decoder/cuda_splatting.py:64-87) so that the batched
entry (`pf3plat_b200.render`) can be fed the same arguments as the reference
function.  On CUDA tensors `make_view_batch` is ONE kernel (gs_view_batch,
csrc/gs_cameras.cu) instead of ~50 small tensor ops with two cuSOLVER inversions;
on CPU tensors (how the tests prepare the oracle's inputs) it is the same
arithmetic in plain tensor code.  The drop-in `GaussianRasterizer` does not need
this module: there the reference's own glue runs unmodified and hands us finished
matrices.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


def get_fov(intrinsics: torch.Tensor) -> torch.Tensor:
    """(B,3,3) normalised intrinsics -> (B,2) full fov (x,y) in radians
    (/root/reference/src/geometry/projection.py:233-247)."""
    inv = torch.linalg.inv(intrinsics)

    def ray(v):
        v = torch.tensor(v, dtype=intrinsics.dtype, device=intrinsics.device)
        d = torch.einsum("bij,j->bi", inv, v)
        return d / d.norm(dim=-1, keepdim=True)

    fov_x = (ray([0.0, 0.5, 1.0]) * ray([1.0, 0.5, 1.0])).sum(-1).acos()
    fov_y = (ray([0.5, 0.0, 1.0]) * ray([0.5, 1.0, 1.0])).sum(-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near, far, fov_x, fov_y) -> torch.Tensor:
    """z in [0,1] perspective matrix, principal point centred
    (/root/reference/src/model/decoder/cuda_splatting.py:17-44)."""
    tx, ty = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    top, right = ty * near, tx * near
    bottom, left = -top, -right
    (b,) = near.shape
    m = torch.zeros((b, 4, 4), dtype=near.dtype, device=near.device)
    m[:, 0, 0] = 2 * near / (right - left)
    m[:, 1, 1] = 2 * near / (top - bottom)
    m[:, 0, 2] = (right + left) / (right - left)
    m[:, 1, 2] = (top + bottom) / (top - bottom)
    m[:, 3, 2] = 1
    m[:, 2, 2] = far / (far - near)
    m[:, 2, 3] = -(far * near) / (far - near)
    return m


@dataclass
class ViewBatch:
    """Per-view camera block in the layout the C-ABI takes (include/gsplat_b200.h: GsView arrays)."""
    viewmatrix: torch.Tensor   # (V,4,4) transposed world->camera
    projmatrix: torch.Tensor   # (V,4,4) transposed full projection
    campos: torch.Tensor       # (V,3)
    tanfov: torch.Tensor       # (V,2)
    scale: torch.Tensor        # (V,) the 1/near rescale applied to means/covariances (1 if not scale-invariant)


def make_view_batch(extrinsics, intrinsics, near, far, scale_invariant: bool = True) -> ViewBatch:
    """cuda_splatting.py:64-87 for all views at once (no per-view `.item()` host syncs)."""
    if extrinsics.is_cuda:
        return _make_view_batch_cuda(extrinsics, intrinsics, near, far, scale_invariant)
    if scale_invariant:
        scale = 1.0 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
        near = near * scale
        far = far * scale
    else:
        scale = torch.ones_like(near)
    fov = get_fov(intrinsics)
    fov_x, fov_y = fov.unbind(dim=-1)
    proj = get_projection_matrix(near, far, fov_x, fov_y).transpose(-1, -2)
    view = torch.linalg.inv(extrinsics).transpose(-1, -2)
    full = view @ proj
    tanfov = torch.stack(((0.5 * fov_x).tan(), (0.5 * fov_y).tan()), dim=-1)
    return ViewBatch(view.contiguous(), full.contiguous(), extrinsics[:, :3, 3].contiguous(), tanfov.contiguous(),
                     scale.contiguous())


def _make_view_batch_cuda(extrinsics, intrinsics, near, far, scale_invariant: bool) -> ViewBatch:
    import ctypes

    from . import _capi
    dev = extrinsics.device
    B = extrinsics.shape[0]
    f = lambda t: t.detach().to(dev, torch.float32).contiguous()
    ext, intr, nr, fr = f(extrinsics), f(intrinsics), f(near), f(far)
    view = torch.empty(B, 4, 4, device=dev)
    proj = torch.empty(B, 4, 4, device=dev)
    campos = torch.empty(B, 3, device=dev)
    tanfov = torch.empty(B, 2, device=dev)
    scale = torch.empty(B, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gs_view_batch(B, int(bool(scale_invariant)), p(ext), p(intr), p(nr), p(fr), p(view), p(proj),
                                              p(campos), p(tanfov), p(scale),
                                              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return ViewBatch(view, proj, campos, tanfov, scale)
