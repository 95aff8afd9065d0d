"""Per-call diagnostics of the PF3plat-shaped batched decoder call (2 scenes x 3 views x 131072 Gaussians, colour + depth):
wall time, library stage times, binning state."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200.rasterizer import last_stats, set_profiling, stage_ms  # noqa: E402
from pf3plat_b200.render import decoder_forward  # noqa: E402
from pf3plat_b200.synthetic import make_scene, make_target  # noqa: E402

b, v, P, hw = 2, 3, 131072, (256, 256)
dev = torch.device("cuda:0")
scs = [make_scene(P, v, *hw, seed=30 + k).to(dev) for k in range(b)]
st = lambda n: torch.stack([getattr(s, n) for s in scs])
means, cov, sh, opac, ext, intr = st("means"), st("covariances"), st("harmonics"), st("opacities"), st("extrinsics"), st("intrinsics")
near = torch.full((b, v), 1.0, device=dev)
far = torch.full((b, v), 100.0, device=dev)
bg = torch.zeros(3, device=dev)
target = make_target(b * v, *hw).to(dev).reshape(b, v, 3, *hw)
if os.environ.get("GS_REF_FIRST"):   # what scripts/bench_decoder.py does before: the reference's per-view call pattern
    from tests.ref_callsite import render_depth_like_reference, render_like_reference
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    rep = lambda t: t.repeat_interleave(v, dim=0)
    for grad in (False, True):
        for _ in range(6):
            L = [t.clone().requires_grad_(grad) for t in (means, cov, sh, opac)]
            c = render_like_reference(flat(ext), flat(intr), flat(near), flat(far), hw, bg[None].expand(b * v, 3), *[rep(t) for t in L])
            d = render_depth_like_reference(flat(ext), flat(intr), flat(near), flat(far), hw, rep(L[0]), rep(L[1]), rep(L[3]))
            if grad:
                (((c.reshape(b, v, 3, *hw) - target) ** 2).mean() + 1e-3 * d.mean()).backward()
    torch.cuda.synchronize()
    print("reference pattern done", last_stats(dev))
set_profiling(bool(int(os.environ.get("GS_PROFILE", "1"))), dev)
for grad in (False, True):
    for k in range(8):
        L = [t.clone().requires_grad_(grad) for t in (means, cov, sh, opac)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c, d = decoder_forward(*L, ext, intr, near, far, hw, bg, depth_mode="depth")
        t1 = time.perf_counter()
        tb = time.perf_counter()
        if grad:
            (((c - target) ** 2).mean() + 1e-3 * d.mean()).backward()
        tb2 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        s = last_stats(dev)
        print(f"grad={grad} call {k}: host-return {1e3 * (t1 - t0):.3f} ms, backward-host {1e3 * (tb2 - tb):.3f} ms, total {1e3 * (t2 - t0):.3f} ms, speculative {s['speculative']}, "
              f"redos {s['overflow_redos']}, D {s['num_rendered']}, stages {{k_: round(v_, 3) for k_, v_ in stage_ms(dev).items() if v_ > 0}}".replace("{{", "{").replace("}}", "}"),
              {k_: round(v_, 3) for k_, v_ in stage_ms(dev).items() if v_ > 0})
