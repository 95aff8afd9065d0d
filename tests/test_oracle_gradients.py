"""Pins the C oracle's hand-written backward: (1) against torch.autograd on an independent vectorised
restatement of the forward (oracle/torch_oracle.py), in fp64; (2) against central finite differences of its own
fp64 forward.  Tolerance 1e-5 relative to each tensor's max-abs (the upstream backward replaces 1/det^2 by
1/(det^2+1e-7), a deliberate ~1e-6-relative deviation from the exact derivative)."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle
from oracle.gs_oracle import OracleRender
from pf3plat_b200.synthetic import make_scene, make_target
from tests.util import view_args


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _scene(P=300, hw=48, seed=3, views=3):
    return make_scene(P, views, hw, hw, seed=seed, d_sh=25)


@pytest.mark.parametrize("mode", ["sh_cov", "rgb_cov", "sh_scalerot", "rgb_cov_depth"])
def test_c_backward_matches_autograd(mode):
    sc = _scene()
    v = 1
    st, kw = view_args(sc, v, use_sh=mode.startswith("sh"))
    if "scalerot" in mode:
        kw.pop("cov3D_precomp")
        kw["scales"] = sc.scales.numpy().astype(np.float64)
        kw["rotations"] = sc.rotations.numpy().astype(np.float64)
    with_depth = mode.endswith("depth")
    r = OracleRender(st, dtype=np.float64, with_depth=with_depth, **kw)
    H, W = sc.image_shape
    target = make_target(1, H, W)[0].double().numpy()
    dL = 2 * (r.color - target) / target.size
    dLd = (np.cos(np.arange(H * W).reshape(H, W)) * 1e-3) if with_depth else None
    g = r.backward(dL, dLd)

    tk = {k: torch.tensor(np.asarray(a, np.float64), requires_grad=True) for k, a in kw.items()}
    m2d = torch.zeros(sc.means.shape[0], 3, dtype=torch.float64, requires_grad=True)
    out = torch_oracle.render(st, means2D=m2d, with_depth=with_depth, **tk)
    color = out[0]
    assert np.abs(color.detach().numpy() - r.color).max() < 1e-9
    assert (out[1].numpy() == r.radii).all()
    loss = (color * torch.tensor(dL)).sum()
    if with_depth:
        assert np.abs(out[2].detach().numpy() - r.depth).max() < 1e-8
        loss = loss + (out[2] * torch.tensor(dLd)).sum()
    loss.backward()
    assert r.num_rendered > 100
    for name, t in tk.items():
        key = name
        ga = t.grad.numpy().reshape(np.asarray(g[key]).shape)
        assert _relerr(g[key], ga) < 1e-5, (name, _relerr(g[key], ga))
    assert _relerr(g["means2D"], m2d.grad.numpy()) < 1e-5


def test_c_backward_matches_finite_differences():
    sc = _scene(P=60, hw=32, seed=5, views=2)
    st, kw = view_args(sc, 0, use_sh=True)
    kw = {k: np.asarray(a, np.float64) for k, a in kw.items()}
    rng = np.random.default_rng(0)
    dL = rng.standard_normal((3, 32, 32))

    def loss(**over):
        k2 = dict(kw)
        k2.update(over)
        return float((OracleRender(st, dtype=np.float64, frag_rel=0, **k2).color * dL).sum())

    g = OracleRender(st, dtype=np.float64, **kw).backward(dL)
    for name in ["means3D", "opacities", "cov3D_precomp", "shs"]:
        base = kw[name]
        flat_g = np.asarray(g[name]).reshape(-1)
        idx = rng.choice(base.size, size=12, replace=False)
        scale = np.abs(flat_g).max()
        for i in idx:
            eps = 1e-6 * max(1.0, abs(base.reshape(-1)[i]))
            p, m = base.copy().reshape(-1), base.copy().reshape(-1)
            p[i] += eps
            m[i] -= eps
            fd = (loss(**{name: p.reshape(base.shape)}) - loss(**{name: m.reshape(base.shape)})) / (2 * eps)
            assert abs(fd - flat_g[i]) < 2e-4 * scale + 1e-9, (name, i, fd, flat_g[i])


def test_fp32_oracle_agrees_with_fp64_outside_fragile_pixels():
    sc = make_scene(10_000, 1, 256, 256, seed=0)       # BASELINE.json configs[0] (C1)
    st, kw = view_args(sc, 0)
    r32 = OracleRender(st, dtype=np.float32, **kw)
    r64 = OracleRender(st, dtype=np.float64, **kw)
    frag = r32.px_fragile | r64.px_fragile
    err = np.abs(r32.color.astype(np.float64) - r64.color).max(axis=0)
    assert frag.mean() < 0.02
    assert err[~frag].max() < 1e-4          # north_star tolerance, abs RGB
    same = r32.radii == r64.radii
    assert (same | r32.geom_fragile | r64.geom_fragile).all()
