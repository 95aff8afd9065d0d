"""Host-side cost of one rasterize_batch call (C2, forward, no_grad): cProfile over 200 calls + wall clock of the
enqueue alone (no synchronisation inside the loop except the library's own verdict wait)."""
import cProfile, io, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pf3plat_b200.cameras import make_view_batch
from pf3plat_b200.rasterizer import BatchSettings, rasterize_batch
from pf3plat_b200.synthetic import make_scene
dev = torch.device("cuda:0")
P, V, HW = 500_000, 8, 256
sc = make_scene(P, V, HW, HW, seed=0).to(dev)
vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
bs = BatchSettings(image_height=HW, image_width=HW, viewmatrix=vb.viewmatrix, projmatrix=vb.projmatrix, campos=vb.campos,
                   bg=sc.background, sh_degree=4, tanfov=vb.tanfov)
c = sc.covariances
args = (sc.means[None].contiguous(), sc.opacities[None].contiguous())
kw = dict(shs=sc.harmonics.permute(0, 2, 1).contiguous()[None],
          cov3D_precomp=torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)[None].contiguous())
def step():
    with torch.no_grad():
        rasterize_batch(bs, *args, **kw)
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host loop {1e3*(t1-t0)/200:.4f} ms/call, with final sync {1e3*(t2-t0)/200:.4f} ms/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
