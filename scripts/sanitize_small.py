"""Small forward+backward over every code path (SH / colours, cov / scale-rot, depth, fast / fallback binning) for
compute-sanitizer:  compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200._capi import (GS_TUNE_BWD_OCC4, GS_TUNE_BWD_V1, GS_TUNE_FORCE_RADIX_BINNING, GS_TUNE_FWD_WS,  # noqa: E402
                                GS_TUNE_PBWD_2PHASE, GS_TUNE_STRATA_MERGE_SORT)
from pf3plat_b200.cameras import make_view_batch  # noqa: E402
from pf3plat_b200.rasterizer import BatchSettings, rasterize_batch  # noqa: E402
from pf3plat_b200.synthetic import make_scene  # noqa: E402

dev = torch.device("cuda:0")
EXTRA = int(os.environ.get("GS_SAN_TUNING", "0"))   # OR-ed into every case (e.g. GS_TUNE_FWD_WS = 32)
for (P, V, hw, tuning, depth, sr, sh) in [(3001, 3, (40, 56), 0, True, False, True),
                                          (777, 1, (16, 16), GS_TUNE_FORCE_RADIX_BINNING, False, True, True),
                                          (5000, 2, (33, 70), 0, True, False, False),
                                          (130, 10, (64, 64), 0, False, False, True),
                                          # round-2 A/B variants: warp-specialised forward + round-1 backward compositor, merge-sorted strata, two-phase
                                          # preprocess backward, 4-CTA backward compositor
                                          (2500, 2, (48, 48), GS_TUNE_FWD_WS | GS_TUNE_BWD_V1 | GS_TUNE_STRATA_MERGE_SORT, True, False, True),
                                          (2500, 2, (48, 48), GS_TUNE_PBWD_2PHASE | GS_TUNE_BWD_OCC4, False, True, True)]:
    sc = make_scene(P, V, *hw, seed=P).to(dev)
    vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
    bs = BatchSettings(image_height=hw[0], image_width=hw[1], viewmatrix=vb.viewmatrix, projmatrix=vb.projmatrix,
                       campos=vb.campos, bg=sc.background, sh_degree=4, tanfov=vb.tanfov, with_depth=depth, tuning=tuning | EXTRA)
    means = sc.means[None].clone().requires_grad_(True)
    opac = sc.opacities[None].clone().requires_grad_(True)
    kw = {}
    if sh:
        kw["shs"] = sc.harmonics.permute(0, 2, 1).contiguous()[None].requires_grad_(True)
    else:
        kw["colors_precomp"] = sc.harmonics[:, :, 0][None].expand(V, P, 3).contiguous().requires_grad_(True)
    if sr:
        kw["scales"] = sc.scales[None].clone().requires_grad_(True)
        kw["rotations"] = sc.rotations[None].clone().requires_grad_(True)
    else:
        c = sc.covariances
        kw["cov3D_precomp"] = torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)[None].requires_grad_(True)
    for rep in range(4):  # exact path (learns capacities + strata), strata trial, strata on their own capacities
        means.grad = None
        out = rasterize_batch(bs, means, opac, **kw)
        loss = out[0].square().mean() + (out[2].mean() if depth else 0)
        loss.backward()
        torch.cuda.synchronize()
        assert torch.isfinite(means.grad).all()
    print("ok", P, V, hw, tuning, depth, sr, sh)

# ---- per-(view, tile) strata: a pixel-aligned cloud (the per-view trial overflows, the redo learns per-tile boundaries) ----
from pf3plat_b200.render import render_views  # noqa: E402
from pf3plat_b200.rasterizer import last_stats  # noqa: E402
from pf3plat_b200.synthetic import make_pixel_aligned_scene  # noqa: E402

sc = make_pixel_aligned_scene(64, 64, 2, seed=3).to(dev)
states = []
for rep in range(6):
    m = sc.means[None].clone().requires_grad_(True)
    col = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sc.background, m, sc.covariances[None],
                       sc.harmonics[None], sc.opacities[None])
    col.square().mean().backward()
    torch.cuda.synchronize()
    states.append(last_stats(dev)["speculative"])
print("ok pixel-aligned, speculative per call:", states)

# ---- gs_render_host with pinned buffers: zero-copy feed (k_sh_colour pulling out of host memory + geometry-only preprocess),
# the fused pull, and the copy-engine path; one scene and two scenes ----
import ctypes  # noqa: E402

from pf3plat_b200 import _capi, rasterizer  # noqa: E402

for (S, P, V, hw) in [(1, 4100, 3, (40, 56)), (2, 1001, 4, (32, 32))]:
    scs = [make_scene(P, V // S, *hw, seed=50 + k) for k in range(S)]
    vbs = [make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far) for sc in scs]
    cat = lambda f: torch.cat([f(k) for k in range(S)]).contiguous().float().pin_memory()
    cov6 = lambda c: torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)
    host = {"means3D": cat(lambda k: scs[k].means), "opacities": cat(lambda k: scs[k].opacities),
            "shs": cat(lambda k: scs[k].harmonics.permute(0, 2, 1)), "cov3D_precomp": cat(lambda k: cov6(scs[k].covariances)),
            "viewmatrix": cat(lambda k: vbs[k].viewmatrix), "projmatrix": cat(lambda k: vbs[k].projmatrix),
            "campos": cat(lambda k: vbs[k].campos), "bg": cat(lambda k: scs[k].background), "tanfov": cat(lambda k: vbs[k].tanfov)}
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
    ref = None
    for tuning in (0, _capi.GS_TUNE_NO_SPLIT_COLOUR, _capi.GS_TUNE_NO_ZERO_COPY):
        cfg.tuning = tuning | EXTRA
        for rep in range(3):
            _capi.check(_capi.lib().gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gout),
                                                   torch.cuda.current_stream(dev).cuda_stream))
            assert torch.isfinite(color).all()
            ref = color.clone() if ref is None else ref
            assert torch.equal(color, ref), (S, tuning)
    print("ok host entry", S, P, V, hw)

# ---- the rows next to the rasterizer: fused adapter (forward + backward), PSNR / SSIM ----
from pf3plat_b200.adapter import GaussianAdapter, GaussianAdapterCfg  # noqa: E402
from pf3plat_b200.metrics import compute_psnr, compute_ssim  # noqa: E402

g = torch.Generator().manual_seed(0)
for (b, v, r, deg) in [(1, 2, 300, 4), (2, 1, 129, 2), (1, 3, 1, 0)]:
    d_in = 7 + 3 * (deg + 1) ** 2
    ext = torch.eye(4).repeat(b, v, 1, 1).reshape(b, v, 1, 4, 4).to(dev)
    intr = torch.tensor([[0.9, 0, 0.5], [0, 0.9, 0.5], [0, 0, 1.0]]).repeat(b, v, 1, 1).reshape(b, v, 1, 3, 3).to(dev)
    raw = torch.randn(b, v, r, d_in, generator=g).to(dev).requires_grad_(True)
    dep = (1 + torch.rand(b, v, r, generator=g)).to(dev).requires_grad_(True)
    xy = torch.rand(b, v, r, 2, generator=g).to(dev)
    opa = torch.rand(b, v, r, generator=g).to(dev)
    ad = GaussianAdapter(GaussianAdapterCfg(0.5, 15.0, deg)).to(dev)
    out = ad(ext, intr, xy, dep, opa, raw, (16, 16))
    (out.means.sum() + out.covariances.sum() + out.harmonics.sum() + out.scales.sum()).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(raw.grad).all() and torch.isfinite(dep.grad).all()
    print("ok adapter", b, v, r, deg)
a, bimg = torch.rand(2, 3, 45, 37, generator=g).to(dev), torch.rand(2, 3, 45, 37, generator=g).to(dev)
print("ok metrics", compute_psnr(a, bimg).tolist(), compute_ssim(a, bimg).tolist())
