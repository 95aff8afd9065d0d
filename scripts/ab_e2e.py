"""A/B of the end-to-end host-buffer entry (gs_render_host) on the C2 workload inside ONE process: alternates the number
of pieces the SH block is fed in (GsConfig.tuning bits 8..11; 1 = one plain copy on the launch stream) so that box-to-box
PCIe differences cancel.  Wall clock per call (the call synchronises)."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf3plat_b200 import _capi, rasterizer  # noqa: E402
from pf3plat_b200.cameras import make_view_batch  # noqa: E402
from pf3plat_b200.synthetic import make_scene  # noqa: E402

P, V, HW = 500_000, 8, 256
dev = torch.device("cuda:0")
sc = make_scene(P, V, HW, HW, seed=0)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
c = sc.covariances
host = {"means3D": sc.means, "opacities": sc.opacities, "shs": sc.harmonics.permute(0, 2, 1).contiguous(),
        "cov3D_precomp": torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1),
        "viewmatrix": vb.viewmatrix, "projmatrix": vb.projmatrix, "campos": vb.campos, "bg": sc.background,
        "tanfov": vb.tanfov}
host = {k: v.contiguous().float().pin_memory() for k, v in host.items()}
cfg = _capi.GsConfig()
cfg.P, cfg.S, cfg.V, cfg.M, cfg.sh_degree = P, 1, V, 25, 4
cfg.image_height = cfg.image_width = HW
cfg.scale_modifier = 1.0
for k in ("viewmatrix", "projmatrix", "campos", "bg", "tanfov"):
    setattr(cfg, k, host[k].data_ptr())
gin = _capi.GsInputs(means3D=host["means3D"].data_ptr(), opacities=host["opacities"].data_ptr(),
                     shs=host["shs"].data_ptr(), cov3D_precomp=host["cov3D_precomp"].data_ptr())
color = torch.empty(V, 3, HW, HW).pin_memory()
radii = torch.empty(V, P, dtype=torch.int32).pin_memory()
gout = _capi.GsOutputs(color=color.data_ptr(), radii=radii.data_ptr(), depth=None)
torch.zeros(1, device=dev)
ctx = rasterizer.current_context(dev)
stream = torch.cuda.current_stream(dev).cuda_stream
L = _capi.lib()


def call():
    _capi.check(L.gs_render_host(ctx, ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gout), stream))


# AB_TUNING: comma-separated full tuning words instead of piece counts (e.g. "0,32768" = images copied back by the copy
# engine vs written directly into the pinned output buffer by the compositor)
if os.environ.get("AB_TUNING"):
    variants = [int(x) for x in os.environ["AB_TUNING"].split(",")]
    word = lambda n: n
else:
    variants = [int(x) for x in os.environ.get("AB_PIECES", "1,2,4,6,8").split(",")]
    word = lambda n: n << _capi.GS_TUNE_FEED_PIECES_SHIFT
for _ in range(5):
    call()
res = {n: [] for n in variants}
for rep in range(int(os.environ.get("AB_REPS", 5))):
    for n in variants:
        cfg.tuning = word(n)
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            call()
        res[n].append((time.perf_counter() - t0) / 20 * 1e3)
if os.environ.get("AB_STAGES"):   # per-stage device times of each variant (events recorded by the library)
    rasterizer.set_profiling(True, dev)
    for n in variants:
        cfg.tuning = word(n)
        acc = {}
        for _ in range(6):
            call()
            for k_, v_ in rasterizer.stage_ms(dev).items():
                acc.setdefault(k_, []).append(v_)
        print(n, {k_: round(sorted(v_)[len(v_) // 2], 4) for k_, v_ in acc.items()})
    rasterizer.set_profiling(False, dev)
for n in variants:
    print(n, " ".join(f"{x:.3f}" for x in res[n]), f"median {sorted(res[n])[len(res[n]) // 2]:.3f} ms  min {min(res[n]):.3f} ms")
