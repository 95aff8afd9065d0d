"""PF3plat-shaped decoder timing (BASELINE.json configs[4] geometry without the encoder): b scenes x v views,
131072 Gaussians per scene (2 x 256 x 256 pixel-aligned Gaussians), 256x256, colour + depth.
  (1) the reference's call pattern: per-view GaussianRasterizer calls on v-fold repeated Gaussians + a second
      pass for depth (tests/ref_callsite.py restates cuda_splatting.py / decoder_splatting_cuda.py);
  (2) pf3plat_b200.render.decoder_forward: one batched call, depth fused.
Prints one JSON line; both run OUR kernels -- this measures what the batched entry buys."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200.render import decoder_forward  # noqa: E402
from pf3plat_b200.synthetic import make_scene, make_target  # noqa: E402
from tests.ref_callsite import render_depth_like_reference, render_like_reference  # noqa: E402

b, v, P, hw = int(os.environ.get("GS_B", 2)), int(os.environ.get("GS_V", 3)), 131072, (256, 256)
steps = int(os.environ.get("GS_STEPS", 10))
dev = torch.device("cuda:0")
scs = [make_scene(P, v, *hw, seed=30 + k).to(dev) for k in range(b)]
st = lambda n: torch.stack([getattr(s, n) for s in scs])
means, cov, sh, opac, ext, intr = st("means"), st("covariances"), st("harmonics"), st("opacities"), st("extrinsics"), st("intrinsics")
near = torch.full((b, v), 1.0, device=dev)
far = torch.full((b, v), 100.0, device=dev)
bg = torch.zeros(3, device=dev)
target = make_target(b * v, *hw).to(dev).reshape(b, v, 3, *hw)
flat = lambda t: t.reshape(b * v, *t.shape[2:])
rep = lambda t: t.repeat_interleave(v, dim=0)


def ref_fwd(grad):
    L = [t.clone().requires_grad_(grad) for t in (means, cov, sh, opac)]
    c = render_like_reference(flat(ext), flat(intr), flat(near), flat(far), hw, bg[None].expand(b * v, 3), *[rep(t) for t in L])
    d = render_depth_like_reference(flat(ext), flat(intr), flat(near), flat(far), hw, rep(L[0]), rep(L[1]), rep(L[3]))
    if grad:
        (((c.reshape(b, v, 3, *hw) - target) ** 2).mean() + 1e-3 * d.mean()).backward()
    return c


def ours_fwd(grad):
    L = [t.clone().requires_grad_(grad) for t in (means, cov, sh, opac)]
    c, d = decoder_forward(*L, ext, intr, near, far, hw, bg, depth_mode="depth")
    if grad:
        (((c - target) ** 2).mean() + 1e-3 * d.mean()).backward()
    return c


def timeit(fn, grad):
    from pf3plat_b200.rasterizer import last_stats
    for _ in range(3):
        fn(grad)
    torch.cuda.synchronize()
    per = []
    t0 = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        fn(grad)
        if os.environ.get("GS_VERBOSE"):
            torch.cuda.synchronize()
            per.append((round(1e3 * (time.perf_counter() - t1), 3), last_stats(dev)["speculative"], last_stats(dev)["overflow_redos"]))
    torch.cuda.synchronize()
    if per:
        sys.stderr.write(f"{fn.__name__} grad={grad}: {per}\n")
    return 1e3 * (time.perf_counter() - t0) / steps


out = {"config": f"{b} scenes x {v} views, {P} Gaussians/scene, 256x256, colour+depth", "steps": steps,
       "reference_call_pattern_ms": {"fwd": timeit(ref_fwd, False), "fwd_bwd": timeit(ref_fwd, True)},
       "batched_decoder_forward_ms": {"fwd": timeit(ours_fwd, False), "fwd_bwd": timeit(ours_fwd, True)}}
out["speedup"] = {k: out["reference_call_pattern_ms"][k] / out["batched_decoder_forward_ms"][k] for k in ("fwd", "fwd_bwd")}
print(json.dumps(out))
