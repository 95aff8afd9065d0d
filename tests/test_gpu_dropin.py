"""GPU tests of the decoder-facing surface: the reference's call pattern (tests/ref_callsite.py restates
cuda_splatting.py line by line) against the batched entry that replaces DecoderSplattingCUDA.forward
(/root/reference/src/model/decoder/decoder_splatting_cuda.py:35-91)."""
import numpy as np
import pytest
import torch

from pf3plat_b200.synthetic import make_scene, make_target
from tests.ref_callsite import decoder_like_reference, render_depth_like_reference, render_like_reference

pytestmark = pytest.mark.gpu


def _scenes(dev, b=2, v=3, P=3000, hw=(48, 64), near=1.0):
    scs = [make_scene(P, v, hw[0], hw[1], seed=20 + k).to(dev) for k in range(b)]
    st = lambda name: torch.stack([getattr(s, name) for s in scs])
    ext, intr = st("extrinsics"), st("intrinsics")
    nr = torch.full((b, v), near, device=dev)
    fr = torch.full((b, v), 100.0 * near, device=dev)
    return scs, st("means"), st("covariances"), st("harmonics"), st("opacities"), ext, intr, nr, fr


@pytest.mark.parametrize("near", [1.0, 0.5])
def test_decoder_forward_matches_the_reference_call_pattern(near):
    from pf3plat_b200.render import decoder_forward
    dev = torch.device("cuda:0")
    b, v, hw = 2, 3, (48, 64)
    scs, means, cov, sh, opac, ext, intr, nr, fr = _scenes(dev, b, v, hw=hw, near=near)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    color, depth = decoder_forward(means, cov, sh, opac, ext, intr, nr, fr, hw, bg, depth_mode="depth")
    assert color.shape == (b, v, 3, *hw) and depth.shape == (b, v, *hw)
    # the reference decoder's call pattern (v-fold repeated Gaussians, per-view op calls, second pass for depth),
    # restated in tests/ref_callsite.py and pinned to the reference's recorded calls by tests/test_camera_cpu.py
    ref_c, ref_d = decoder_like_reference(means, cov, sh, opac, ext, intr, nr, fr, hw, bg, depth_mode="depth")
    # same kernels, same per-view arithmetic: the only difference is where the 1/near rescale is applied
    assert (color - ref_c).abs().max() <= 1e-4
    assert (depth - ref_d).abs().max() <= 2e-3 * float(ref_d.abs().max())


@pytest.mark.parametrize("mode", ["disparity", "relative_disparity", "log"])
def test_other_depth_modes(mode):
    from pf3plat_b200.render import render_depth
    dev = torch.device("cuda:0")
    b, v, hw = 1, 2, (32, 48)
    scs, means, cov, sh, opac, ext, intr, nr, fr = _scenes(dev, b, v, P=1500, hw=hw)
    got = render_depth(means, cov, opac, ext, intr, nr, fr, hw, mode=mode)
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    rep = lambda t: t.repeat_interleave(v, dim=0)
    ref = render_depth_like_reference(flat(ext), flat(intr), flat(nr), flat(fr), hw, rep(means), rep(cov), rep(opac),
                                      mode=mode)
    assert (got.reshape(b * v, *hw) - ref).abs().max() <= 1e-4 * max(1.0, float(ref.abs().max()))


def test_gradients_flow_to_the_encoder_side_tensors():
    from pf3plat_b200.render import decoder_forward
    dev = torch.device("cuda:0")
    b, v, hw = 2, 2, (32, 32)
    scs, means, cov, sh, opac, ext, intr, nr, fr = _scenes(dev, b, v, P=1000, hw=hw)
    leaves = [t.clone().requires_grad_(True) for t in (means, cov, sh, opac)]
    color, depth = decoder_forward(*leaves, ext, intr, nr, fr, hw, torch.zeros(3, device=dev), depth_mode="depth")
    target = make_target(b * v, *hw).to(dev).reshape(b, v, 3, *hw)
    (((color - target) ** 2).mean() + 1e-3 * depth.mean()).backward()
    # reference pattern with autograd through the per-view loop and the v-fold repeat
    leaves_r = [t.clone().requires_grad_(True) for t in (means, cov, sh, opac)]
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    rep = lambda t: t.repeat_interleave(v, dim=0)
    rc = render_like_reference(flat(ext), flat(intr), flat(nr), flat(fr), hw, torch.zeros(b * v, 3, device=dev),
                               *[rep(t) for t in leaves_r])
    rd = render_depth_like_reference(flat(ext), flat(intr), flat(nr), flat(fr), hw, rep(leaves_r[0]), rep(leaves_r[1]),
                                     rep(leaves_r[3]))
    (((rc.reshape(b, v, 3, *hw) - target) ** 2).mean() + 1e-3 * rd.mean()).backward()
    for a, r in zip(leaves, leaves_r):
        assert a.grad is not None and torch.isfinite(a.grad).all()
        err = (a.grad - r.grad).abs().max() / r.grad.abs().max()
        assert err <= 1e-3, float(err)


def test_depth_gradient_reaches_the_extrinsics_like_the_reference():
    """PF3plat trains with depth_mode "depth" on PREDICTED poses (config/main.yaml:50): the reference's depth pass builds
    z = (extrinsics^-1 @ mean).z in torch (cuda_splatting.py:239-242), so a depth loss reaches the extrinsics.  The fused
    colour+depth pass cannot carry that gradient (cameras enter the kernels detached), so decoder_forward must take the
    two-pass route whenever the extrinsics require grad -- and give the reference's gradient."""
    from pf3plat_b200.render import decoder_forward
    dev = torch.device("cuda:0")
    b, v, hw = 1, 2, (32, 32)
    scs, means, cov, sh, opac, ext, intr, nr, fr = _scenes(dev, b, v, P=1500, hw=hw)
    bg = torch.zeros(3, device=dev)
    w = torch.rand(b, v, *hw, device=dev)
    ext_a = ext.clone().requires_grad_(True)
    means_a = means.clone().requires_grad_(True)
    color, depth = decoder_forward(means_a, cov, sh, opac, ext_a, intr, nr, fr, hw, bg, depth_mode="depth")
    (depth * w).sum().backward()
    ext_r = ext.clone().requires_grad_(True)
    means_r = means.clone().requires_grad_(True)
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    rep = lambda t: t.repeat_interleave(v, dim=0)
    rd = render_depth_like_reference(flat(ext_r), flat(intr), flat(nr), flat(fr), hw, rep(means_r), rep(cov), rep(opac))
    (rd.reshape(b, v, *hw) * w).sum().backward()
    assert ext_a.grad is not None and float(ext_r.grad.abs().max()) > 0
    err = (ext_a.grad - ext_r.grad).abs().max() / ext_r.grad.abs().max()
    assert err <= 1e-3, float(err)
    err_m = (means_a.grad - means_r.grad).abs().max() / means_r.grad.abs().max()
    assert err_m <= 1e-3, float(err_m)
    # without a pose gradient the fused single pass is used and gives the same depth
    with torch.no_grad():
        _, depth_fused = decoder_forward(means, cov, sh, opac, ext, intr, nr, fr, hw, bg, depth_mode="depth")
    assert (depth_fused - depth.detach()).abs().max() <= 2e-3 * float(depth.abs().max())


def test_inplace_update_between_forward_and_backward_raises_and_pool_trims():
    from pf3plat_b200.rasterizer import last_stats, trim_memory
    from pf3plat_b200.render import render_views
    dev = torch.device("cuda:0")
    sc = make_scene(2000, 1, 32, 32, seed=6).to(dev)
    means = sc.means[None].clone().requires_grad_(True)
    color = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (32, 32), sc.background, means,
                         sc.covariances[None], sc.harmonics[None], sc.opacities[None])
    with torch.no_grad():
        means.mul_(1.0)      # an optimizer step / clamp_ on a leaf between forward and backward
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        color.sum().backward()
    del color
    trim_memory(dev)         # the library's private pool gives its cached blocks back; next call still works
    again = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (32, 32), sc.background, sc.means[None],
                         sc.covariances[None], sc.harmonics[None], sc.opacities[None])
    assert torch.isfinite(again).all() and last_stats(dev)["num_rendered"] > 0


def test_orthographic_style_settings_with_tensor_tanfov():
    """render_cuda_orthographic passes tanfovx/tanfovy as 0-dim CUDA tensors (cuda_splatting.py:195-196)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda:0")
    sc = make_scene(500, 1, 32, 32, seed=5).to(dev)
    from pf3plat_b200.cameras import make_view_batch
    vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
    common = dict(image_height=32, image_width=32, bg=sc.background[0], scale_modifier=1.0,
                  viewmatrix=vb.viewmatrix[0], projmatrix=vb.projmatrix[0], sh_degree=4, campos=vb.campos[0],
                  prefiltered=False, debug=False)
    row, col = torch.triu_indices(3, 3)
    args = dict(means3D=sc.means, means2D=torch.zeros_like(sc.means), shs=sc.harmonics.permute(0, 2, 1).contiguous(),
                opacities=sc.opacities[:, None], cov3D_precomp=sc.covariances[:, row, col])
    a, _ = GaussianRasterizer(GaussianRasterizationSettings(tanfovx=vb.tanfov[0, 0], tanfovy=vb.tanfov[0, 1], **common))(**args)
    b_, _ = GaussianRasterizer(GaussianRasterizationSettings(tanfovx=float(vb.tanfov[0, 0]), tanfovy=float(vb.tanfov[0, 1]),
                                                            **common))(**args)
    assert torch.equal(a, b_)
    vis = GaussianRasterizer(GaussianRasterizationSettings(tanfovx=0.58, tanfovy=0.58, **common)).markVisible(sc.means)
    assert vis.dtype == torch.bool and vis.all()      # every synthetic Gaussian sits at z >= 1.5


@pytest.mark.parametrize("P", [20000, 70001])
def test_host_buffer_entry_matches_device_entry(P):
    """gs_render_host (C ABI with HOST pointers: the e2e path of bench.py) against the torch-facing device path.
    70001 Gaussians: the SH block (21 MB) goes over in pieces with preprocess running piece by piece behind them;
    called three times so that the exact, trial and stratified binning paths all see the pieced feed."""
    import ctypes
    from pf3plat_b200 import _capi, rasterizer
    from pf3plat_b200.cameras import make_view_batch
    from pf3plat_b200.rasterizer import BatchSettings, rasterize_batch
    dev = torch.device("cuda:0")
    V, hw = 3, (64, 80)
    sc = make_scene(P, V, *hw, seed=9)
    vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
    c = sc.covariances
    host = {"means3D": sc.means, "opacities": sc.opacities, "shs": sc.harmonics.permute(0, 2, 1).contiguous(),
            "cov3D_precomp": torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1),
            "viewmatrix": vb.viewmatrix, "projmatrix": vb.projmatrix, "campos": vb.campos, "bg": sc.background,
            "tanfov": vb.tanfov}
    host = {k: v.contiguous().float() for k, v in host.items()}          # pageable host memory is fine too
    cfg = _capi.GsConfig()
    cfg.P, cfg.S, cfg.V, cfg.M, cfg.sh_degree = P, 1, V, 25, 4
    cfg.image_height, cfg.image_width, cfg.scale_modifier = hw[0], hw[1], 1.0
    cfg.flags = _capi.GS_FLAG_DEPTH
    for k in ("viewmatrix", "projmatrix", "campos", "bg", "tanfov"):
        setattr(cfg, k, host[k].data_ptr())
    gin = _capi.GsInputs(means3D=host["means3D"].data_ptr(), opacities=host["opacities"].data_ptr(),
                         shs=host["shs"].data_ptr(), cov3D_precomp=host["cov3D_precomp"].data_ptr())
    color = torch.empty(V, 3, *hw)
    radii = torch.empty(V, P, dtype=torch.int32)
    depth = torch.empty(V, *hw)
    gout = _capi.GsOutputs(color=color.data_ptr(), radii=radii.data_ptr(), depth=depth.data_ptr())
    ctx = rasterizer.current_context(dev)
    d = {k: v.to(dev) for k, v in host.items()}
    bs = BatchSettings(image_height=hw[0], image_width=hw[1], viewmatrix=d["viewmatrix"], projmatrix=d["projmatrix"],
                       campos=d["campos"], bg=d["bg"], sh_degree=4, tanfov=d["tanfov"], with_depth=True)
    c2, r2, d2 = rasterize_batch(bs, d["means3D"][None], d["opacities"][None], shs=d["shs"][None],
                                 cov3D_precomp=d["cov3D_precomp"][None])
    for _ in range(3):
        color.zero_(); radii.zero_(); depth.zero_()
        _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gout),
                                               torch.cuda.current_stream(dev).cuda_stream))
        assert torch.equal(color, c2.cpu()) and torch.equal(radii, r2.cpu()) and torch.equal(depth, d2.cpu())
    # PINNED output buffers, and the experiment in which the compositor writes the images straight into them
    # (GS_TUNE_DIRECT_OUTPUT; default: device-to-host copy afterwards) -- same bytes either way
    pc, pr, pd = color.clone().pin_memory(), radii.clone().pin_memory(), depth.clone().pin_memory()
    gpin = _capi.GsOutputs(color=pc.data_ptr(), radii=pr.data_ptr(), depth=pd.data_ptr())
    for tuning in (0, _capi.GS_TUNE_DIRECT_OUTPUT):
        cfg.tuning = tuning
        pc.zero_(); pr.zero_(); pd.zero_()
        _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gpin),
                                               torch.cuda.current_stream(dev).cuda_stream))
        assert torch.equal(pc, c2.cpu()) and torch.equal(pr, r2.cpu()) and torch.equal(pd, d2.cpu())
    cfg.tuning = 0
    # PINNED input buffers (default): geometry-only preprocess + tile sort on the launch stream while k_sh_colour pulls the
    # SH pieces it wants straight out of the host buffer on a second stream (zero-copy feed, split pipeline);
    # GS_TUNE_NO_SPLIT_COLOUR = one fused preprocess doing the pull; GS_TUNE_NO_ZERO_COPY = the copy engine's pieced upload;
    # GS_TUNE_PRE_SH_RAW16 = the pull's staging on the device copy -- same bytes out every way.  A view of the pinned block
    # that starts one row in (only 4-byte aligned) must take the upload path by itself.
    pin = {k: v.pin_memory() for k, v in host.items()}
    gin_pin = _capi.GsInputs(means3D=pin["means3D"].data_ptr(), opacities=pin["opacities"].data_ptr(),
                             shs=pin["shs"].data_ptr(), cov3D_precomp=pin["cov3D_precomp"].data_ptr())
    for tuning in (0, _capi.GS_TUNE_NO_SPLIT_COLOUR, _capi.GS_TUNE_NO_ZERO_COPY, _capi.GS_TUNE_NO_ZERO_COPY | _capi.GS_TUNE_PRE_SH_RAW16):
        cfg.tuning = tuning
        for _ in range(3):   # (exact, trial and stratified binning under every feed)
            pc.zero_(); pr.zero_(); pd.zero_()
            _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin_pin), ctypes.byref(gpin),
                                                   torch.cuda.current_stream(dev).cuda_stream))
            assert torch.equal(pc, c2.cpu()) and torch.equal(pr, r2.cpu()) and torch.equal(pd, d2.cpu()), tuning
    cfg.tuning = 0
    shifted = torch.empty(P + 1, 25, 3).pin_memory()
    shifted[1:] = host["shs"]
    gin_sh = _capi.GsInputs(means3D=pin["means3D"].data_ptr(), opacities=pin["opacities"].data_ptr(),
                            shs=shifted[1:].data_ptr(), cov3D_precomp=pin["cov3D_precomp"].data_ptr())
    assert shifted[1:].data_ptr() % 16 != 0
    pc.zero_()
    _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin_sh), ctypes.byref(gpin),
                                           torch.cuda.current_stream(dev).cuda_stream))
    assert torch.equal(pc, c2.cpu())
    # invalid argument combinations come back as the reference op's exceptions, not crashes
    bad = _capi.GsInputs(means3D=host["means3D"].data_ptr(), opacities=host["opacities"].data_ptr())
    with pytest.raises(ValueError, match="SHs or precomputed colors"):
        _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(bad), ctypes.byref(gout),
                                               torch.cuda.current_stream(dev).cuda_stream))


def test_host_buffer_entry_with_two_scenes_and_pinned_inputs():
    """gs_render_host, S = 2 scenes x 2 views each, pinned buffers: k_sh_colour's scene indexing (zero-copy feed) against the
    device entry and against the copy-engine path."""
    import ctypes
    from pf3plat_b200 import _capi, rasterizer
    from pf3plat_b200.cameras import make_view_batch
    from pf3plat_b200.rasterizer import BatchSettings, rasterize_batch
    dev = torch.device("cuda:0")
    S, P, V, hw = 2, 9004, 4, (48, 64)     # P * 300 bytes is a multiple of 16: the second scene's block stays aligned
    scs = [make_scene(P, V // S, *hw, seed=70 + k) for k in range(S)]
    vbs = [make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far) for sc in scs]
    cat = lambda f: torch.cat([f(k) for k in range(S)]).contiguous().float().pin_memory()
    cov6 = lambda c: torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)
    host = {"means3D": cat(lambda k: scs[k].means), "opacities": cat(lambda k: scs[k].opacities),
            "shs": cat(lambda k: scs[k].harmonics.permute(0, 2, 1)), "cov3D_precomp": cat(lambda k: cov6(scs[k].covariances)),
            "viewmatrix": cat(lambda k: vbs[k].viewmatrix), "projmatrix": cat(lambda k: vbs[k].projmatrix),
            "campos": cat(lambda k: vbs[k].campos), "bg": cat(lambda k: scs[k].background), "tanfov": cat(lambda k: vbs[k].tanfov)}
    d = {k: v.to(dev) for k, v in host.items()}
    bs = BatchSettings(image_height=hw[0], image_width=hw[1], viewmatrix=d["viewmatrix"], projmatrix=d["projmatrix"],
                       campos=d["campos"], bg=d["bg"], sh_degree=4, tanfov=d["tanfov"])
    c2, r2 = rasterize_batch(bs, d["means3D"].reshape(S, P, 3), d["opacities"].reshape(S, P), shs=d["shs"].reshape(S, P, 25, 3),
                             cov3D_precomp=d["cov3D_precomp"].reshape(S, P, 6))
    cfg = _capi.GsConfig()
    cfg.P, cfg.S, cfg.V, cfg.M, cfg.sh_degree = P, S, V, 25, 4
    cfg.image_height, cfg.image_width, cfg.scale_modifier = hw[0], hw[1], 1.0
    for k in ("viewmatrix", "projmatrix", "campos", "bg", "tanfov"):
        setattr(cfg, k, host[k].data_ptr())
    gin = _capi.GsInputs(means3D=host["means3D"].data_ptr(), opacities=host["opacities"].data_ptr(), shs=host["shs"].data_ptr(),
                         cov3D_precomp=host["cov3D_precomp"].data_ptr())
    color = torch.empty(V, 3, *hw).pin_memory()
    radii = torch.empty(V, P, dtype=torch.int32).pin_memory()
    gout = _capi.GsOutputs(color=color.data_ptr(), radii=radii.data_ptr(), depth=None)
    ctx = rasterizer.current_context(dev)
    for tuning in (0, _capi.GS_TUNE_NO_SPLIT_COLOUR, _capi.GS_TUNE_NO_ZERO_COPY):
        cfg.tuning = tuning
        for _ in range(3):
            color.zero_(); radii.zero_()
            _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gout),
                                                   torch.cuda.current_stream(dev).cuda_stream))
            assert torch.equal(color, c2.cpu()) and torch.equal(radii, r2.cpu()), tuning


def test_psnr_matches_the_reference_formula():
    """compute_psnr against the reference's formula (src/evaluation/metrics.py:11-19) evaluated in fp64."""
    from pf3plat_b200.metrics import compute_psnr
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for shape in [(5, 3, 37, 53), (2, 3, 256, 256), (1, 1, 1, 3)]:
        gt = torch.rand(shape, generator=g) * 1.4 - 0.2          # exercises the clip on both sides
        pr = torch.rand(shape, generator=g) * 1.4 - 0.2
        ref = -10 * ((gt.double().clip(0, 1) - pr.double().clip(0, 1)) ** 2).flatten(1).mean(1).log10()
        got = compute_psnr(gt.to(dev), pr.to(dev)).cpu().double()
        assert torch.allclose(got, ref, atol=1e-4, rtol=0), (got, ref)
        assert torch.equal(compute_psnr(gt.to(dev), pr.to(dev)).cpu().double(), got)     # deterministic


@pytest.mark.parametrize("shape", [(4, 3, 256, 256), (2, 3, 37, 53), (1, 1, 11, 11), (3, 2, 43, 11)])
def test_ssim_matches_the_restated_skimage_algorithm(shape):
    """compute_ssim (gs_ssim) vs oracle/ssim_oracle.py (scipy's gaussian_filter, fp64).  Tolerance 2e-5 absolute: the
    kernel filters in fp32 (skimage, on float32 input, does too)."""
    from oracle import ssim_oracle
    from pf3plat_b200.metrics import compute_ssim
    g = torch.Generator().manual_seed(shape[-1])
    gt = torch.rand(*shape, generator=g)
    ys, xs = torch.meshgrid(torch.linspace(0, 6.28, shape[2]), torch.linspace(0, 6.28, shape[3]), indexing="ij")
    gt = 0.5 * gt + 0.25 * (1 + torch.sin(3 * xs) * torch.cos(2 * ys))          # structure + noise
    pred = (gt + 0.1 * torch.randn(*shape, generator=g)).clamp(0, 1)
    pred[0] = gt[0]                                                              # identical pair -> exactly 1
    got = compute_ssim(gt.cuda(), pred.cuda()).cpu().double().numpy()
    want = ssim_oracle.compute_ssim(gt.numpy(), pred.numpy())
    np.testing.assert_allclose(got, want, atol=2e-5, rtol=0)
    assert abs(got[0] - 1.0) < 1e-6


def test_ssim_rejects_images_smaller_than_the_window():
    from pf3plat_b200.metrics import compute_ssim
    with pytest.raises(ValueError, match="win_size exceeds image extent"):
        compute_ssim(torch.zeros(1, 3, 10, 64, device="cuda"), torch.zeros(1, 3, 10, 64, device="cuda"))


@pytest.mark.parametrize("scale_invariant", [True, False])
def test_camera_glue_kernel_matches_the_tensor_restatement(scale_invariant):
    """gs_view_batch (one kernel, fp64 inside) against the plain-tensor restatement of cuda_splatting.py:64-87 that
    tests/ref_callsite.py and the CPU path of make_view_batch use; general (non-rigid) extrinsics included, since the
    reference calls a general .inverse().  Tolerance: fp32 rounding of the tensor path (~1e-6 relative)."""
    from pf3plat_b200.cameras import make_view_batch
    g = torch.Generator().manual_seed(4)
    B = 37
    ext = torch.eye(4).repeat(B, 1, 1)
    ext[:, :3, :3] = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    ext[:, :3, 3] = torch.randn(B, 3, generator=g)
    ext[B // 2:, :3, :3] *= 1.0 + 0.1 * torch.rand(B - B // 2, 1, 1, generator=g)          # not orthonormal
    intr = torch.eye(3).repeat(B, 1, 1)
    intr[:, 0, 0] = 0.5 + torch.rand(B, generator=g)
    intr[:, 1, 1] = 0.5 + torch.rand(B, generator=g)
    intr[:, 0, 2] = 0.5 + 0.05 * torch.randn(B, generator=g)
    intr[:, 1, 2] = 0.5 + 0.05 * torch.randn(B, generator=g)
    near = 0.2 + torch.rand(B, generator=g)
    far = 50 + 100 * torch.rand(B, generator=g)
    cpu = make_view_batch(ext.double(), intr.double(), near.double(), far.double(), scale_invariant)   # fp64 truth
    gpu = make_view_batch(ext.cuda(), intr.cuda(), near.cuda(), far.cuda(), scale_invariant)
    for name in ("viewmatrix", "projmatrix", "campos", "tanfov", "scale"):
        a, b = getattr(gpu, name).cpu().double(), getattr(cpu, name).double()
        assert a.shape == b.shape, name
        assert (a - b).abs().max() <= 2e-6 * max(1.0, float(b.abs().max())), (name, float((a - b).abs().max()))


def test_c5_chain_runs_end_to_end():
    """scripts/c5_chain.py at a small size: stand-in encoder -> fused adapter -> batched decoder -> MSE -> backward ->
    optimizer step -> PSNR/SSIM, all on the device; gradients reach the encoder's weights and are finite."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, C5_HW="48", C5_STEPS="2", C5_TARGETS="2")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "c5_chain.py")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["gradients_finite"] and line["n_gpus"] == 1 and line["ms_per_step"] > 0
    assert 0.0 < line["loss_last"] < 1.0 and -1.0 <= line["ssim"] <= 1.0


@pytest.mark.parametrize("tag,scale_invariant", [("si", True), ("raw", False)])
def test_camera_glue_kernel_matches_what_the_reference_hands_its_rasterizer(tag, scale_invariant):
    """gs_view_batch against the settings the reference's own render_cuda produced (tests/golden/camera_glue.npz)."""
    from pf3plat_b200.cameras import make_view_batch
    from tests.test_camera_cpu import GOLDEN, check_against_golden
    z = np.load(GOLDEN)
    t = lambda k: torch.from_numpy(z[k]).cuda()
    vb = make_view_batch(t("extrinsics"), t("intrinsics"), t("near"), t("far"), scale_invariant)
    check_against_golden(vb, z, tag, 3e-6)
