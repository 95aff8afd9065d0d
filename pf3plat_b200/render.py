"""Batched render entry: the decoder-facing functions of the reference with the per-view Python loop, the v-fold
`repeat` of the Gaussians and the per-view host syncs removed.

Replaces, behind the same argument meaning:
  * `render_cuda`        /root/reference/src/model/decoder/cuda_splatting.py:47-127
  * `render_depth_cuda`  /root/reference/src/model/decoder/cuda_splatting.py:226-269 (mode "depth" fused as a
                         4th composited channel; other modes render the fake colour like the reference)
  * the body of `DecoderSplattingCUDA.forward`  /root/reference/src/model/decoder/decoder_splatting_cuda.py:35-67
All views of all scenes go through ONE gs_forward call (one host sync per batch instead of three per view).
"""
from __future__ import annotations

from math import isqrt
from typing import Optional

import torch

from .cameras import make_view_batch
from .rasterizer import BatchSettings, rasterize_batch

_TRIU: dict = {}


def _cov6(cov: torch.Tensor) -> torch.Tensor:
    """(..., 3, 3) -> (..., 6) upper triangle in the order the call site uses (cuda_splatting.py:115,123).
    One gather forward and one index_add backward (the obvious stack of six slices costs ~25 small kernels per step)."""
    idx = _TRIU.get(cov.device)
    if idx is None:
        idx = _TRIU[cov.device] = torch.tensor([0, 1, 2, 4, 5, 8], device=cov.device)
    return torch.index_select(cov.flatten(-2), -1, idx)


def render_views(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                 gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant: bool = True,
                 use_sh: bool = True, with_depth: bool = False, sh_layout: str = "g_xyz_n"):
    """extrinsics (B,4,4) c2w, intrinsics (B,3,3), near/far (B,), background (B,3); Gaussians are per SCENE and NOT
    repeated per view: means (S,G,3), covariances (S,G,3,3) or (S,G,6), sh (S,G,3,d_sh) [or (S,G,d_sh,3) with
    sh_layout="g_n_xyz"], opacities (S,G); B % S == 0 and view b shows scene b // (B//S).
    Returns color (B,3,h,w) and, if with_depth, depth (B,h,w) = sum_i alpha_i T_i z_i."""
    assert use_sh or gaussian_sh_coefficients.shape[-1 if sh_layout == "g_xyz_n" else -2] == 1
    h, w = image_shape
    vb = make_view_batch(extrinsics, intrinsics, near, far, scale_invariant)
    if sh_layout == "g_xyz_n":
        n = gaussian_sh_coefficients.shape[-1]
        shs = gaussian_sh_coefficients.transpose(-1, -2)  # (S,G,n,3); made contiguous by the operator
    else:
        n = gaussian_sh_coefficients.shape[-2]
        shs = gaussian_sh_coefficients
    degree = isqrt(n) - 1
    cov6 = gaussian_covariances if gaussian_covariances.shape[-1] == 6 and gaussian_covariances.dim() == 3 \
        else _cov6(gaussian_covariances)
    bs = BatchSettings(image_height=h, image_width=w, viewmatrix=vb.viewmatrix, projmatrix=vb.projmatrix,
                       campos=vb.campos, bg=background_color, sh_degree=degree, tanfov=vb.tanfov,
                       view_scale=vb.scale if scale_invariant else None, with_depth=with_depth)
    B = extrinsics.shape[0]
    S = gaussian_means.shape[0]
    if use_sh:
        out = rasterize_batch(bs, gaussian_means, gaussian_opacities, shs=shs, cov3D_precomp=cov6)
    else:
        colors = shs[..., 0, :]                                  # (S,G,3)
        colors = colors.repeat_interleave(B // S, dim=0) if S != B else colors
        out = rasterize_batch(bs, gaussian_means, gaussian_opacities, colors_precomp=colors, cov3D_precomp=cov6)
    if with_depth:
        return out[0], out[2]
    return out[0]


def render_cuda(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant: bool = True,
                use_sh: bool = True):
    """Signature-compatible with the reference's render_cuda (Gaussians already repeated per view: batch == views)."""
    return render_views(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                        gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant, use_sh)


def decoder_forward(means, covariances, harmonics, opacities, extrinsics, intrinsics, near, far, image_shape,
                    background_color, depth_mode: Optional[str] = None):
    """`DecoderSplattingCUDA.forward` (decoder_splatting_cuda.py:35-67) without the v-fold repeat: Gaussians
    (b,g,...), cameras (b,v,...).  Returns (color (b,v,3,h,w), depth (b,v,h,w) or None).  depth_mode "depth" is
    fused into the colour pass; the reference's other modes are served by `render_depth`.

    Gradients: like the reference's colour pass, the fused pass gives the CAMERAS no gradient (the view matrices enter
    the rasterizer detached on both sides, cuda_splatting.py:85-87 / make_view_batch) and routes dL/ddepth to the means
    only.  The reference's depth pass, however, builds its fake colour z = (extrinsics^-1 @ mean).z in torch
    (cuda_splatting.py:239-242), so its depth IS differentiable w.r.t. the extrinsics -- PF3plat trains with
    depth_mode "depth" on predicted poses.  Whenever the extrinsics require grad, "depth" therefore takes the
    reference's two-pass route (`render_depth`), which carries that gradient."""
    b, v = extrinsics.shape[:2]
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    bg = background_color.reshape(1, 3).expand(b * v, 3)
    pose_grad = depth_mode == "depth" and extrinsics.requires_grad and torch.is_grad_enabled()
    if (depth_mode is None or depth_mode == "depth") and not pose_grad:
        out = render_views(flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape, bg, means,
                           covariances, harmonics, opacities, with_depth=depth_mode is not None)
        if depth_mode is None:
            return out.reshape(b, v, *out.shape[1:]), None
        color, depth = out
        # the reference renders depth in the 1/near-rescaled scene (render_depth_cuda -> render_cuda with
        # scale_invariant=True) but takes the fake colour from the UNSCALED camera-space z
        # (cuda_splatting.py:239-242): undo the rescale of z
        depth = depth * flat(near)[:, None, None]
        return color.reshape(b, v, *color.shape[1:]), depth.reshape(b, v, *depth.shape[1:])
    color = render_views(flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape, bg, means,
                         covariances, harmonics, opacities)
    depth = render_depth(means, covariances, opacities, extrinsics, intrinsics, near, far, image_shape, depth_mode)
    return color.reshape(b, v, *color.shape[1:]), depth


def render_depth(means, covariances, opacities, extrinsics, intrinsics, near, far, image_shape, mode: str = "depth"):
    """`render_depth_cuda` (cuda_splatting.py:226-269) for (b,g,...) Gaussians and (b,v,...) cameras."""
    b, v = extrinsics.shape[:2]
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    ext, nr, fr = flat(extrinsics), flat(near), flat(far)
    w2c = torch.linalg.inv(ext)                                                   # (B,4,4)
    m = means.repeat_interleave(v, dim=0) if v > 1 else means                     # (B,g,3) view of scene per view
    z = torch.einsum("bj,bgj->bg", w2c[:, 2, :3], m) + w2c[:, 2, 3:4]
    if mode == "disparity":
        z = 1 / z
    elif mode == "relative_disparity":
        eps = 1e-10
        dn, df, d = 1 / (nr[:, None] + eps), 1 / (fr[:, None] + eps), 1 / (z + eps)
        z = 1 - (d - df) / (dn - df + eps)
    elif mode == "log":
        z = z.minimum(nr[:, None]).maximum(fr[:, None]).log()
    fake = z[..., None, None].expand(-1, -1, 3, 1)                                # (B,g,3,1)
    # fake colours are per VIEW, so this path needs Gaussians indexed per view: S == B
    cov = covariances.repeat_interleave(v, dim=0) if v > 1 else covariances
    opa = opacities.repeat_interleave(v, dim=0) if v > 1 else opacities
    out = render_views(ext, flat(intrinsics), nr, fr, image_shape, torch.zeros((b * v, 3), device=ext.device), m,
                       cov, fake, opa, use_sh=False)
    return out.mean(dim=1).reshape(b, v, *image_shape)
