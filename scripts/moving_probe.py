"""Why is the forward slower when the cloud changes every step?  Renders 12 variants of the C2 cloud in turn and prints the
per-stage times and the un-instrumented ms/step for: the static cloud, jitter only, re-draw only, both; and for both with
the variants' tensors pre-touched."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200.cameras import make_view_batch  # noqa: E402
from pf3plat_b200.rasterizer import BatchSettings, last_stats, rasterize_batch, set_profiling, stage_ms  # noqa: E402
from pf3plat_b200.synthetic import make_scene  # noqa: E402

P, V, HW = 500_000, 8, 256
dev = torch.device("cuda:0")
sc = make_scene(P, V, HW, HW, seed=0)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
bs = BatchSettings(image_height=HW, image_width=HW, viewmatrix=vb.viewmatrix.to(dev), projmatrix=vb.projmatrix.to(dev),
                   campos=vb.campos.to(dev), bg=sc.background.to(dev), sh_degree=4, tanfov=vb.tanfov.to(dev))
c = sc.covariances
cov6 = torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)[None].to(dev)
shs = sc.harmonics.permute(0, 2, 1).contiguous()[None].to(dev)
opac = sc.opacities[None].to(dev)
g = torch.Generator().manual_seed(7)
px = 2.0 * (0.5 / 0.86) / HW


def variants(jitter, redraw, n=12):
    out = []
    for _ in range(n):
        m = sc.means.clone()
        if jitter:
            m = m + torch.randn(P, 3, generator=g) * (sc.means[:, 2:3] * px) * torch.tensor([1.0, 1.0, 0.0])
        if redraw:
            r = torch.rand(P, generator=g) < 0.05
            m[r] = sc.means[torch.randperm(P, generator=g)[: int(r.sum())]]
        out.append(m.reshape(1, P, 3).contiguous().to(dev))
    return out


def run(name, vs):
    with torch.no_grad():
        for k in range(6):
            rasterize_batch(bs, vs[k % len(vs)], opac, shs=shs, cov3D_precomp=cov6)
        set_profiling(True, dev)
        acc, spec = {}, []
        for k in range(12):
            rasterize_batch(bs, vs[k % len(vs)], opac, shs=shs, cov3D_precomp=cov6)
            torch.cuda.synchronize()
            for kk, v in stage_ms(dev).items():
                acc.setdefault(kk, []).append(v)
            spec.append(last_stats(dev)["speculative"])
        set_profiling(False, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for k in range(24):
            rasterize_batch(bs, vs[k % len(vs)], opac, shs=shs, cov3D_precomp=cov6)
        e1.record()
        torch.cuda.synchronize()
    st = {k: round(statistics.median(v), 4) for k, v in acc.items() if statistics.median(v) > 0}
    print(f"{name:28s} {e0.elapsed_time(e1) / 24:.4f} ms/step  stages {st}  speculative {sorted(set(spec))}  D {last_stats(dev)['num_rendered']}")


run("static", [sc.means.reshape(1, P, 3).to(dev)])
run("static x12 copies", [sc.means.reshape(1, P, 3).to(dev).clone() for _ in range(12)])
run("jitter only", variants(True, False))
run("redraw only", variants(False, True))
run("jitter + redraw", variants(True, True))
run("static again", [sc.means.reshape(1, P, 3).to(dev)])
