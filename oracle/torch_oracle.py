"""Independent vectorised PyTorch restatement of the rasterizer (fp64 by default), gradients by autograd.

TEST INFRASTRUCTURE ONLY (same rules as gs_oracle.c).  Purpose: pin the C
oracle's hand-written analytic backward against torch.autograd on the same
forward semantics, for small inputs.  PARITY UNPINNED
with respect to the reference's external CUDA extension (absent here).

Differences from a naive autograd graph that are needed to follow upstream:
  * alpha = min(0.99, o*G) passes its gradient straight through the clamp;
  * the guard-band clamp of t.x/t.y masks d/dt.x (resp. y) and treats the
    clamped value as a constant w.r.t. t.z;
  * skip / termination decisions are masks (no gradient).
"""
from __future__ import annotations

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def _sh_rgb(deg, shs, dirs):
    x, y, z = dirs.unbind(-1)
    x, y, z = x[:, None], y[:, None], z[:, None]
    r = SH_C0 * shs[:, 0]
    if deg > 0:
        r = r - SH_C1 * y * shs[:, 1] + SH_C1 * z * shs[:, 2] - SH_C1 * x * shs[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = (r + SH_C2[0] * xy * shs[:, 4] + SH_C2[1] * yz * shs[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * shs[:, 6]
             + SH_C2[3] * xz * shs[:, 7] + SH_C2[4] * (xx - yy) * shs[:, 8])
    if deg > 2:
        r = (r + SH_C3[0] * y * (3 * xx - yy) * shs[:, 9] + SH_C3[1] * xy * z * shs[:, 10]
             + SH_C3[2] * y * (4 * zz - xx - yy) * shs[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12]
             + SH_C3[4] * x * (4 * zz - xx - yy) * shs[:, 13] + SH_C3[5] * z * (xx - yy) * shs[:, 14]
             + SH_C3[6] * x * (xx - 3 * yy) * shs[:, 15])
    return torch.clamp_min(r + 0.5, 0.0)


def render(settings, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
           cov3D_precomp=None, means2D=None, with_depth=False, dtype=torch.float64):
    """settings: oracle.gs_oracle.OracleSettings.  Tensor args may require grad.  Returns (color, radii[, depth])."""
    s = settings
    H, W = int(s.image_height), int(s.image_width)
    t = lambda a: torch.as_tensor(a, dtype=dtype)
    view, proj, campos, bg = t(s.viewmatrix).reshape(4, 4), t(s.projmatrix).reshape(4, 4), t(s.campos), t(s.bg)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dtype)
    hom = torch.cat([means3D, ones], dim=-1)
    p_view = hom @ view          # row-vector convention: matrices arrive transposed
    p_hom = hom @ proj
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    if means2D is not None:
        ndc = ndc + means2D[:, :2]
    if cov3D_precomp is not None:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], -1).reshape(P, 3, 3)
    else:
        r, x, y, z = rotations.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(P, 3, 3)
        sv = s.scale_modifier * scales
        Sigma = R @ torch.diag_embed(sv * sv) @ R.transpose(-1, -2)
    tz = p_view[:, 2]
    limx, limy = s.guard_band * s.tanfovx, s.guard_band * s.tanfovy
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    inx, iny = (txtz >= -limx) & (txtz <= limx), (tytz >= -limy) & (tytz <= limy)
    tx = torch.where(inx, p_view[:, 0], (txtz.clamp(-limx, limx) * tz).detach())
    ty = torch.where(iny, p_view[:, 1], (tytz.clamp(-limy, limy) * tz).detach())
    fx, fy = W / (2 * s.tanfovx), H / (2 * s.tanfovy)
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz)], -1).reshape(P, 2, 3)
    Rv = view[:3, :3].T   # standard rotation
    Mm = J @ Rv
    cov = Mm @ Sigma @ Mm.transpose(-1, -2)
    a, b, c_ = cov[:, 0, 0] + s.dilation, cov[:, 0, 1], cov[:, 1, 1] + s.dilation
    det = a * c_ - b * b
    valid = (tz > s.near_cull_z) & (det != 0)
    det_safe = torch.where(det != 0, det, torch.ones_like(det))
    conic = torch.stack([c_ / det_safe, -b / det_safe, a / det_safe], -1)
    mid = 0.5 * (a + c_)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    lam2 = mid - torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(lam, lam2))).detach()
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16
    def trunc_clamp(v, hi):
        return torch.clamp(torch.trunc(v.detach()).to(torch.int64), 0, hi)
    rminx, rminy = trunc_clamp((px - radius) / 16, gx), trunc_clamp((py - radius) / 16, gy)
    rmaxx, rmaxy = trunc_clamp((px + radius + 15) / 16, gx), trunc_clamp((py + radius + 15) / 16, gy)
    valid = valid & ((rmaxx - rminx) * (rmaxy - rminy) > 0)
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        deg = min(int(s.sh_degree), int(s.sh_eval_max_degree))
        d = means3D - campos
        rgb = _sh_rgb(deg, shs, d / d.norm(dim=-1, keepdim=True))
    radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)

    depth32 = tz.detach().to(torch.float32)
    color = torch.zeros(3, H, W, dtype=dtype)
    depth_img = torch.zeros(H, W, dtype=dtype)
    for ty_ in range(gy):
        for tx_ in range(gx):
            sel = valid & (rminx <= tx_) & (tx_ < rmaxx) & (rminy <= ty_) & (ty_ < rmaxy)
            ids = torch.nonzero(sel).squeeze(-1)
            if ids.numel():
                # stable order: (depth fp32 bits, index)
                order = torch.argsort(depth32[ids], stable=True)
                ids = ids[order]
            ys = torch.arange(ty_ * 16, min(ty_ * 16 + 16, H))
            xs = torch.arange(tx_ * 16, min(tx_ * 16 + 16, W))
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            pxf, pyf = xx.reshape(-1).to(dtype), yy.reshape(-1).to(dtype)
            n = pxf.numel()
            if ids.numel() == 0:
                C = torch.zeros(n, 3, dtype=dtype)
                Dp = torch.zeros(n, dtype=dtype)
                Tfin = torch.ones(n, dtype=dtype)
            else:
                dx = px[ids][None, :] - pxf[:, None]
                dy = py[ids][None, :] - pyf[:, None]
                con = conic[ids]
                power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
                G = torch.exp(torch.clamp_max(power, 0.0))
                araw = opacities.reshape(-1)[ids][None, :] * G
                alpha = araw + (torch.clamp_max(araw, 0.99) - araw).detach()
                ok = (power <= 0) & (alpha >= 1.0 / 255.0)
                a_eff = torch.where(ok, alpha, torch.zeros_like(alpha))
                Tincl = torch.cumprod(1.0 - a_eff, dim=1)
                Tbefore = torch.cat([torch.ones(n, 1, dtype=dtype), Tincl[:, :-1]], dim=1)
                stop = ok & (Tincl < 1e-4)
                stopped = torch.cumsum(stop.to(torch.int64), dim=1) > 0   # inclusive: the stopping Gaussian is NOT added
                live = ok & ~stopped
                w = torch.where(live, a_eff * Tbefore, torch.zeros_like(a_eff))
                C = w @ rgb[ids]
                Dp = w @ tz[ids]
                a_live = torch.where(live, a_eff, torch.zeros_like(a_eff))
                Tfin = torch.prod(1.0 - a_live, dim=1)
            out = C + Tfin[:, None] * bg[None, :]
            color[:, yy.reshape(-1), xx.reshape(-1)] = out.T
            depth_img[yy.reshape(-1), xx.reshape(-1)] = Dp
    if with_depth:
        return color, radii, depth_img
    return color, radii
