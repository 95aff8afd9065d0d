"""Randomised consistency check on the GPU: for random shapes (Gaussians, views, image size, scale / depth distortions that
crowd depth strata or blow up footprints), every call of a shape -- exact, strata trial, stratified, per-tile strata, after
overflows -- must give the pixels, radii and depths of the exact-capacity path BIT FOR BIT, forward+backward must give the
gradients of the round-1 kernels within 2e-5, and gs_render_host (pinned: zero-copy split / fused; pageable: copy engine)
the device entry's bytes.  Usage: python scripts/fuzz_paths.py [cases] [seed]"""
import ctypes, os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200 import _capi, rasterizer
from pf3plat_b200.cameras import make_view_batch
from pf3plat_b200.rasterizer import BatchSettings, last_stats, rasterize_batch
from pf3plat_b200.synthetic import make_pixel_aligned_scene, make_scene

dev = torch.device("cuda:0")
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for case in range(cases):
    kind = rng.choice(["plain", "plain", "flat_depth", "big", "aligned", "tiny"])
    V = rng.randint(1, 5)
    hw = (rng.choice([16, 33, 64, 100, 128, 200]), rng.choice([16, 48, 64, 112, 160, 256]))
    if kind == "aligned":
        side = rng.choice([32, 48, 64])
        sc = make_pixel_aligned_scene(side, side, V, seed=case)
        hw = sc.image_shape
    else:
        P = rng.choice([1, 7, 130, 1000, 5000, 20000, 60000]) if kind != "tiny" else rng.randint(1, 40)
        sc = make_scene(P, V, *hw, seed=1000 + case)
        if kind == "flat_depth":      # all Gaussians at (nearly) one depth: strata crowd, ties by index
            sc.means[:, 2] = 5.0 + (torch.arange(sc.means.shape[0]) % 3).float() * 1e-6
        if kind == "big":             # large footprints: long tile lists, radix fallback possible
            sc.covariances.mul_(rng.choice([25.0, 400.0]))
    d = sc.to(dev)
    P = d.means.shape[0]
    vb = make_view_batch(d.extrinsics, d.intrinsics, d.near, d.far)
    c = d.covariances
    cov6 = torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)[None].contiguous()
    shs = d.harmonics.permute(0, 2, 1).contiguous()[None]
    depth = rng.random() < 0.5
    mk = lambda tuning: BatchSettings(image_height=hw[0], image_width=hw[1], viewmatrix=vb.viewmatrix, projmatrix=vb.projmatrix,
                                      campos=vb.campos, bg=d.background, sh_degree=4, tanfov=vb.tanfov, view_scale=vb.scale,
                                      with_depth=depth, tuning=tuning)
    with torch.no_grad():
        ref = [t.clone() for t in rasterize_batch(mk(_capi.GS_TUNE_NO_SPECULATION), d.means[None], d.opacities[None], shs=shs, cov3D_precomp=cov6)]
    states = []
    for rep in range(5):
        with torch.no_grad():
            out = rasterize_batch(mk(0), d.means[None], d.opacities[None], shs=shs, cov3D_precomp=cov6)
        states.append(last_stats(dev)["speculative"])
        for a, b in zip(out, ref):
            if not torch.equal(a, b):
                bad += 1
                print("MISMATCH forward", case, kind, P, V, hw, "rep", rep, states, float((a.float() - b.float()).abs().max()))
    # same shape, different content: the cloud pulled towards the optical axis (central tiles get several times the
    # instances the capacities were learned for -> overflow, exact redo, re-learning), then the original again
    if rng.random() < 0.5 and P > 100:
        m2 = d.means.clone()
        m2[:, :2] *= rng.choice([0.2, 0.5])
        with torch.no_grad():
            ref2 = [t.clone() for t in rasterize_batch(mk(_capi.GS_TUNE_NO_SPECULATION), m2[None], d.opacities[None], shs=shs, cov3D_precomp=cov6)]
            redos0 = last_stats(dev)["overflow_redos"]
            for means_k, ref_k in ((m2, ref2), (m2, ref2), (d.means, ref), (m2, ref2), (d.means, ref)):
                out = rasterize_batch(mk(0), means_k[None], d.opacities[None], shs=shs, cov3D_precomp=cov6)
                for a, b in zip(out, ref_k):
                    if not torch.equal(a, b):
                        bad += 1
                        print("MISMATCH forward after content change", case, kind, P, V, hw)
        overflow_seen = overflow_seen + (last_stats(dev)["overflow_redos"] - redos0) if "overflow_seen" in dir() else last_stats(dev)["overflow_redos"] - redos0
    # gradients: default kernels vs round-1 compositor + two-phase preprocess backward
    g = {}
    for tuning in (0, _capi.GS_TUNE_BWD_V1):
        leaves = [t.clone().requires_grad_(True) for t in (d.means[None], d.opacities[None], shs, cov6)]
        out = rasterize_batch(mk(tuning), leaves[0], leaves[1], shs=leaves[2], cov3D_precomp=leaves[3])
        w = torch.randn(out[0].shape, generator=torch.Generator().manual_seed(case)).to(dev)
        ((out[0] * w).sum() + (out[2].sum() * 0.01 if depth else 0.0)).backward()
        g[tuning] = [t.grad for t in leaves]
    for a, b in zip(g[0], g[_capi.GS_TUNE_BWD_V1]):
        if float((a - b).abs().max()) > (1e-4 if kind == "big" else 2e-5) * float(b.abs().max()) + 1e-12 or not torch.isfinite(a).all():
            bad += 1
            print("MISMATCH backward", case, kind, P, V, hw, float((a - b).abs().max()), float(b.abs().max()))
    # host entry
    host = {"means3D": d.means.cpu(), "opacities": d.opacities.cpu(), "shs": shs[0].cpu(), "cov3D_precomp": cov6[0].cpu(),
            "viewmatrix": vb.viewmatrix.cpu(), "projmatrix": vb.projmatrix.cpu(), "campos": vb.campos.cpu(), "bg": d.background.cpu(),
            "tanfov": vb.tanfov.cpu(), "view_scale": vb.scale.cpu()}
    host = {k: v.contiguous().float() for k, v in host.items()}
    for pinned in (True, False):
        hb = {k: (v.pin_memory() if pinned else v) for k, v in host.items()}
        cfg = _capi.GsConfig()
        cfg.P, cfg.S, cfg.V, cfg.M, cfg.sh_degree = P, 1, V, 25, 4
        cfg.image_height, cfg.image_width, cfg.scale_modifier = hw[0], hw[1], 1.0
        cfg.flags = _capi.GS_FLAG_DEPTH if depth else 0
        for k in ("viewmatrix", "projmatrix", "campos", "bg", "tanfov", "view_scale"):
            setattr(cfg, k, hb[k].data_ptr())
        gin = _capi.GsInputs(means3D=hb["means3D"].data_ptr(), opacities=hb["opacities"].data_ptr(), shs=hb["shs"].data_ptr(),
                             cov3D_precomp=hb["cov3D_precomp"].data_ptr())
        color = torch.empty(V, 3, *hw); radii = torch.empty(V, P, dtype=torch.int32); dep = torch.empty(V, *hw)
        gout = _capi.GsOutputs(color=color.data_ptr(), radii=radii.data_ptr(), depth=dep.data_ptr() if depth else None)
        for tuning in ((0, _capi.GS_TUNE_NO_SPLIT_COLOUR) if pinned else (0,)):
            cfg.tuning = tuning
            _capi.check(_capi.lib().gs_render_host(rasterizer.current_context(dev), ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gout),
                                                   torch.cuda.current_stream(dev).cuda_stream))
            if not (torch.equal(color, ref[0].cpu()) and torch.equal(radii, ref[1].cpu()) and (not depth or torch.equal(dep, ref[2].cpu()))):
                bad += 1
                print("MISMATCH host entry", case, kind, P, V, hw, "pinned", pinned, "tuning", tuning)
    print(f"case {case:3d} {kind:10s} P={P:6d} V={V} hw={hw} depth={int(depth)} speculative per call {states}", flush=True)
# ---- second part: two scenes per call, scales + rotations or covariances, SH or precomputed colours ----
for case in range(cases // 2):
    S = rng.choice([1, 2, 3])
    vps = rng.randint(1, 3)
    V = S * vps
    hw = (rng.choice([24, 64, 96]), rng.choice([32, 64, 144]))
    P = rng.choice([3, 500, 4000, 15000])
    scs = [make_scene(P, vps, *hw, seed=5000 + 10 * case + k).to(dev) for k in range(S)]
    vbs = [make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far) for sc in scs]
    cat = lambda f: torch.cat([f(k) for k in range(S)]).contiguous()
    stack = lambda f: torch.stack([f(k) for k in range(S)]).contiguous()
    use_sr, use_sh, depth = rng.random() < 0.5, rng.random() < 0.6, rng.random() < 0.5
    kw = {}
    if use_sh:
        kw["shs"] = stack(lambda k: scs[k].harmonics.permute(0, 2, 1))
    else:
        kw["colors_precomp"] = cat(lambda k: scs[k].harmonics[:, :, 0][None].expand(vps, P, 3)).abs()
    if use_sr:
        kw["scales"] = stack(lambda k: scs[k].scales)
        kw["rotations"] = stack(lambda k: scs[k].rotations)
    else:
        c6 = lambda c: torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)
        kw["cov3D_precomp"] = stack(lambda k: c6(scs[k].covariances))
    means, opac = stack(lambda k: scs[k].means), stack(lambda k: scs[k].opacities)
    mk = lambda tuning: BatchSettings(image_height=hw[0], image_width=hw[1], viewmatrix=cat(lambda k: vbs[k].viewmatrix),
                                      projmatrix=cat(lambda k: vbs[k].projmatrix), campos=cat(lambda k: vbs[k].campos),
                                      bg=cat(lambda k: scs[k].background), sh_degree=4, tanfov=cat(lambda k: vbs[k].tanfov),
                                      view_scale=cat(lambda k: vbs[k].scale), with_depth=depth, tuning=tuning)
    with torch.no_grad():
        ref = [t.clone() for t in rasterize_batch(mk(_capi.GS_TUNE_NO_SPECULATION), means, opac, **kw)]
        # every scene alone gives its own views of the batched call
        for k in range(S):
            one = BatchSettings(image_height=hw[0], image_width=hw[1], viewmatrix=vbs[k].viewmatrix, projmatrix=vbs[k].projmatrix,
                                campos=vbs[k].campos, bg=scs[k].background, sh_degree=4, tanfov=vbs[k].tanfov, view_scale=vbs[k].scale,
                                with_depth=depth, tuning=_capi.GS_TUNE_NO_SPECULATION)
            kw1 = {n: (t[k * vps:(k + 1) * vps] if n == "colors_precomp" else t[k:k + 1]) for n, t in kw.items()}
            alone = rasterize_batch(one, means[k:k + 1], opac[k:k + 1], **kw1)
            for a, b in zip(alone, ref):
                if not torch.equal(a, b[k * vps:(k + 1) * vps]):
                    bad += 1
                    print("MISMATCH scene alone vs batched", case, S, P, vps, hw, k)
    for rep in range(4):
        with torch.no_grad():
            out = rasterize_batch(mk(0), means, opac, **kw)
        for a, b in zip(out, ref):
            if not torch.equal(a, b):
                bad += 1
                print("MISMATCH forward (scenes)", case, S, P, vps, hw, "rep", rep)
    g = {}
    for tuning in (0, _capi.GS_TUNE_BWD_V1 | _capi.GS_TUNE_PBWD_2PHASE):
        lv = {n: t.clone().requires_grad_(True) for n, t in dict(kw, means=means, opac=opac).items()}
        out = rasterize_batch(mk(tuning), lv["means"], lv["opac"], **{n: lv[n] for n in kw})
        w = torch.randn(out[0].shape, generator=torch.Generator().manual_seed(case)).to(dev)
        ((out[0] * w).sum() + (out[2].sum() * 0.01 if depth else 0.0)).backward()
        g[tuning] = {n: t.grad for n, t in lv.items()}
    for n in g[0]:
        a, b = g[0][n], g[_capi.GS_TUNE_BWD_V1 | _capi.GS_TUNE_PBWD_2PHASE][n]
        if float((a - b).abs().max()) > 2e-5 * float(b.abs().max()) + 1e-12 or not torch.isfinite(a).all():
            bad += 1
            print("MISMATCH backward (scenes)", case, n, S, P, vps, hw, float((a - b).abs().max()), float(b.abs().max()))
    print(f"scenes case {case:3d} S={S} P={P:6d} views/scene={vps} hw={hw} sr={int(use_sr)} sh={int(use_sh)} depth={int(depth)}", flush=True)
print("FUZZ", "FAILED" if bad else "ok", "mismatches", bad, "| overflow redos provoked:", overflow_seen if "overflow_seen" in dir() else 0)
sys.exit(1 if bad else 0)
