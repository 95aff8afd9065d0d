"""Line-by-line restatement of how PF3plat calls the rasterizer (the reference tree is absent on the GPU box, so
the GPU tests cannot import it).  Mirrors /root/reference/src/model/decoder/cuda_splatting.py:47-127
(`render_cuda`): 1/near rescale, SH relayout, fov/projection, and the PER-VIEW loop that builds
GaussianRasterizationSettings with `.item()` floats, a zero `means2D` that requires grad, and
`cov[:, row, col]` from `torch.triu_indices`.  The CPU test
tests/test_capi_cpu.py::test_reference_render_glue_imports_and_reaches_our_operator_unmodified drives the real file."""
from math import isqrt

import torch
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

from pf3plat_b200.cameras import get_fov, get_projection_matrix


def render_like_reference(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                          gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant=True,
                          use_sh=True, return_radii=False, return_means2d=False):
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    if scale_invariant:
        scale = 1 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
        gaussian_covariances = gaussian_covariances * (scale[:, None, None, None] ** 2)
        gaussian_means = gaussian_means * scale[:, None, None]
        near = near * scale
        far = far * scale
    _, _, _, n = gaussian_sh_coefficients.shape
    degree = isqrt(n) - 1
    shs = gaussian_sh_coefficients.permute(0, 1, 3, 2).contiguous()
    b, _, _ = extrinsics.shape
    h, w = image_shape
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    tan_fov_x = (0.5 * fov_x).tan()
    tan_fov_y = (0.5 * fov_y).tan()
    projection_matrix = get_projection_matrix(near, far, fov_x, fov_y).transpose(-1, -2)
    view_matrix = extrinsics.inverse().transpose(-1, -2)
    full_projection = view_matrix @ projection_matrix
    all_images, all_radii, all_m2d = [], [], []
    for i in range(b):
        mean_gradients = torch.zeros_like(gaussian_means[i], requires_grad=True)
        try:
            mean_gradients.retain_grad()
        except Exception:
            pass
        settings = GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=tan_fov_x[i].item(), tanfovy=tan_fov_y[i].item(),
            bg=background_color[i], scale_modifier=1.0, viewmatrix=view_matrix[i], projmatrix=full_projection[i],
            sh_degree=degree, campos=extrinsics[i, :3, 3], prefiltered=False, debug=False)
        rasterizer = GaussianRasterizer(settings)
        row, col = torch.triu_indices(3, 3)
        image, radii = rasterizer(
            means3D=gaussian_means[i], means2D=mean_gradients, shs=shs[i] if use_sh else None,
            colors_precomp=None if use_sh else shs[i, :, 0, :], opacities=gaussian_opacities[i, ..., None],
            cov3D_precomp=gaussian_covariances[i, :, row, col])
        all_images.append(image)
        all_radii.append(radii)
        all_m2d.append(mean_gradients)
    out = torch.stack(all_images)
    extra = []
    if return_radii:
        extra.append(torch.stack(all_radii))
    if return_means2d:
        extra.append(all_m2d)
    return (out, *extra) if extra else out


def render_depth_like_reference(extrinsics, intrinsics, near, far, image_shape, gaussian_means, gaussian_covariances,
                                gaussian_opacities, scale_invariant=True, mode="depth"):
    """Restates render_depth_cuda (/root/reference/src/model/decoder/cuda_splatting.py:226-269): camera-space z (or
    a function of it) rendered as a 3-channel colour with bg = 0, then averaged over the channels."""
    hom = torch.cat([gaussian_means, torch.ones_like(gaussian_means[..., :1])], dim=-1)
    cam = torch.einsum("bij,bgj->bgi", extrinsics.inverse(), hom)
    fake = cam[..., 2]
    if mode == "disparity":
        fake = 1 / fake
    elif mode == "relative_disparity":
        eps = 1e-10
        dn, df, d = 1 / (near[:, None] + eps), 1 / (far[:, None] + eps), 1 / (fake + eps)
        fake = 1 - (d - df) / (dn - df + eps)
    elif mode == "log":
        fake = fake.minimum(near[:, None]).maximum(far[:, None]).log()
    b = fake.shape[0]
    result = render_like_reference(extrinsics, intrinsics, near, far, image_shape,
                                   torch.zeros((b, 3), dtype=fake.dtype, device=fake.device), gaussian_means,
                                   gaussian_covariances, fake[..., None, None].expand(-1, -1, 3, 1), gaussian_opacities,
                                   scale_invariant=scale_invariant, use_sh=False)
    return result.mean(dim=1)


def decoder_like_reference(means, covariances, harmonics, opacities, extrinsics, intrinsics, near, far, image_shape,
                           background_color, depth_mode=None):
    """Restates DecoderSplattingCUDA.forward / .render_depth
    (/root/reference/src/model/decoder/decoder_splatting_cuda.py:35-91): cameras flattened to (b v), the Gaussians of
    each scene REPEATED v times ("b g ... -> (b v) g ..."), one colour render and -- if depth_mode is given -- a second
    render for the depth.  Returns (color (b,v,3,h,w), depth (b,v,h,w) or None)."""
    b, v = extrinsics.shape[:2]
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    rep = lambda t: t.repeat_interleave(v, dim=0)
    color = render_like_reference(flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape,
                                  background_color[None].expand(b * v, 3), rep(means), rep(covariances), rep(harmonics),
                                  rep(opacities))
    color = color.reshape(b, v, *color.shape[1:])
    if depth_mode is None:
        return color, None
    depth = render_depth_like_reference(flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape, rep(means),
                                        rep(covariances), rep(opacities), mode=depth_mode)
    return color, depth.reshape(b, v, *depth.shape[1:])


def orthographic_settings_like_reference(extrinsics, width, height, near, far, image_shape, background_color, d_sh,
                                         fov_degrees: float = 0.1):
    """The camera part of render_cuda_orthographic (/root/reference/src/model/decoder/cuda_splatting.py:154-177, :192-205):
    a fake "orthographic" projection -- tiny field of view, camera moved back along its own z axis.  Returns one dict per
    view with the fields the reference puts into GaussianRasterizationSettings (numpy, for the oracle) plus the tensors."""
    b = extrinsics.shape[0]
    h, w = image_shape
    fov_x = torch.tensor(fov_degrees, device=extrinsics.device).deg2rad()
    tan_fov_x = (0.5 * fov_x).tan()
    distance_to_near = (0.5 * width) / tan_fov_x
    tan_fov_y = 0.5 * height / distance_to_near
    fov_y = (2 * tan_fov_y).atan()
    near = near + distance_to_near
    far = far + distance_to_near
    move_back = torch.eye(4, dtype=torch.float32, device=extrinsics.device)
    move_back[2, 3] = -distance_to_near
    extrinsics = extrinsics @ move_back
    projection_matrix = get_projection_matrix(near, far, fov_x.expand(b), fov_y).transpose(-1, -2)
    view_matrix = extrinsics.inverse().transpose(-1, -2)
    full_projection = view_matrix @ projection_matrix
    out = []
    for i in range(b):
        out.append(dict(tanfovx=tan_fov_x, tanfovy=tan_fov_y[i] if tan_fov_y.dim() else tan_fov_y, bg=background_color[i],
                        viewmatrix_t=view_matrix[i], projmatrix_t=full_projection[i], campos_t=extrinsics[i, :3, 3],
                        sh_degree=isqrt(d_sh) - 1))
    for d in out:
        d["viewmatrix"] = d["viewmatrix_t"].detach().cpu().numpy()
        d["projmatrix"] = d["projmatrix_t"].detach().cpu().numpy()
        d["campos"] = d["campos_t"].detach().cpu().numpy()
        d["tanfovx"], d["tanfovy"] = float(d["tanfovx"]), float(d["tanfovy"])
    return out


def render_orthographic_like_reference(extrinsics, width, height, near, far, image_shape, background_color,
                                       gaussian_means, gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities,
                                       fov_degrees: float = 0.1, use_sh: bool = True):
    """Restates render_cuda_orthographic (cuda_splatting.py:130-220).  Unlike render_cuda it passes the 0-dim TENSOR
    tan_fov_x / tan_fov_y into the settings (:195-196), not `.item()` floats."""
    b = extrinsics.shape[0]
    h, w = image_shape
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    n = gaussian_sh_coefficients.shape[-1]
    shs = gaussian_sh_coefficients.permute(0, 1, 3, 2).contiguous()
    sts = orthographic_settings_like_reference(extrinsics, width, height, near, far, image_shape, background_color, n,
                                               fov_degrees)
    fov_x = torch.tensor(fov_degrees, device=extrinsics.device).deg2rad()
    tan_fov_x = (0.5 * fov_x).tan()
    distance_to_near = (0.5 * width) / tan_fov_x
    tan_fov_y = 0.5 * height / distance_to_near
    all_images = []
    for i in range(b):
        mean_gradients = torch.zeros_like(gaussian_means[i], requires_grad=True)
        try:
            mean_gradients.retain_grad()
        except Exception:
            pass
        settings = GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=tan_fov_x, tanfovy=tan_fov_y, bg=background_color[i],
            scale_modifier=1.0, viewmatrix=sts[i]["viewmatrix_t"], projmatrix=sts[i]["projmatrix_t"],
            sh_degree=sts[i]["sh_degree"], campos=sts[i]["campos_t"], prefiltered=False, debug=False)
        rasterizer = GaussianRasterizer(settings)
        row, col = torch.triu_indices(3, 3)
        image, radii = rasterizer(
            means3D=gaussian_means[i], means2D=mean_gradients, shs=shs[i] if use_sh else None,
            colors_precomp=None if use_sh else shs[i, :, 0, :], opacities=gaussian_opacities[i, ..., None],
            cov3D_precomp=gaussian_covariances[i, :, row, col])
        all_images.append(image)
    return torch.stack(all_images)
