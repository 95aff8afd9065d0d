"""Fused adapter kernels (pf3plat_b200/adapter.py -> gs_adapter_forward/backward through the C ABI) against
 * the golden fixtures produced by the reference's own GaussianAdapter (tests/golden/adapter_*.npz), and
 * the fp64 oracle (oracle/adapter_oracle.py) on larger seeded inputs, including ragged / unaligned block sizes.
Tolerances (fp32 kernels vs fp64 truth): 1e-5 relative to the largest magnitude for outputs, 1e-3 for gradients
(BASELINE.json's gradient tolerance); measured errors are ~1e-6 / ~1e-5."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import adapter_util as au  # noqa: E402

pytestmark = pytest.mark.gpu


def _adapter(sh_degree):
    from pf3plat_b200.adapter import GaussianAdapter, GaussianAdapterCfg
    return GaussianAdapter(GaussianAdapterCfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, sh_degree=sh_degree)).cuda()


def _run_kernels(z, meta, with_d=True):
    dev = torch.device("cuda:0")
    leaves = {k: torch.from_numpy(z["in_" + k]).to(dev).requires_grad_(k != "opacities") for k in au.INPUTS}
    d = torch.from_numpy(z["sh_rotation"]).to(dev) if with_d else None
    out = _adapter(meta["sh_degree"]).forward(leaves["extrinsics"], leaves["intrinsics"], leaves["coordinates"],
                                              leaves["depths"], leaves["opacities"], leaves["raw_gaussians"],
                                              (meta["h"], meta["w"]), sh_rotation=d)
    outs = {k: getattr(out, k) for k in au.OUTPUTS}
    loss = sum((torch.from_numpy(z["w_" + k]).to(dev) * outs[k]).sum() for k in au.OUTPUTS)
    loss.backward()
    return out, {k: v.detach().cpu() for k, v in outs.items()}, {k: leaves[k].grad.cpu() for k in au.GRADS}


@pytest.mark.parametrize("name", au.CASES)
def test_kernels_match_the_reference_fixture(name):
    z, meta = au.load(name)
    out, outs, grads = _run_kernels(z, meta)
    for k in au.OUTPUTS:
        assert outs[k].shape == z["out_" + k].shape, k
        assert au.rel_err(outs[k], z["out_" + k]) < 1e-5, (k, au.rel_err(outs[k], z["out_" + k]))
    for k in au.GRADS:
        assert grads[k].shape == z["grad_" + k].shape, k
        assert au.rel_err(grads[k], z["grad_" + k]) < 1e-3, (k, au.rel_err(grads[k], z["grad_" + k]))
    # opacities are handed through untouched, harmonics are the transposed view of the rasterizer layout
    assert out.harmonics.transpose(-1, -2).is_contiguous()


def test_default_sh_rotation_is_derived_from_the_extrinsics():
    """Without an explicit sh_rotation the module builds the Wigner-D blocks itself (proper rotations) or uses the
    identity (improper ones), as the reference's rotate_sh does."""
    for name in ("adapter_pf3plat", "adapter_improper"):
        z, meta = au.load(name)
        _, outs, _ = _run_kernels(z, meta, with_d=False)
        assert au.rel_err(outs["harmonics"], z["out_harmonics"]) < 1e-5


@pytest.mark.parametrize("b,v,r,deg", [(1, 2, 65536, 4), (2, 3, 1001, 4), (1, 1, 1, 4), (1, 5, 131, 0), (3, 1, 259, 3)])
def test_kernels_match_oracle_on_larger_and_ragged_inputs(b, v, r, deg):
    """65536 = PF3plat's 256x256 context view; 1001 / 131 / 259 / 1 exercise the tail block and the unaligned
    (non-TMA) staging paths (row sizes 328 B and 300 B are only 16-byte aligned at even / 4-aligned offsets)."""
    sys.path.insert(0, au.GOLDEN)
    import make_adapter_golden as mk
    from oracle import adapter_oracle
    from pf3plat_b200.sh_rotation import sh_rotation_blocks
    h, w = 1, r
    inp = mk.make_inputs(7 + r, b, v, h, w, deg)
    d_sh = (deg + 1) ** 2
    dmat = sh_rotation_blocks(inp["extrinsics"][..., :3, :3].reshape(b, v, 3, 3).double(), d_sh)
    g = torch.Generator().manual_seed(5)
    # oracle (fp64, CPU)
    lo = {k: t.double().requires_grad_(k != "opacities") for k, t in inp.items()}
    oo = adapter_oracle.adapter_forward(lo["extrinsics"], lo["intrinsics"], lo["coordinates"], lo["depths"],
                                        lo["opacities"], lo["raw_gaussians"], (h, w), deg, 0.5, 15.0,
                                        sh_rotation=dmat[:, :, None, None, None])
    weights = {k: torch.randn(oo[k].shape, generator=g, dtype=torch.float64) for k in au.OUTPUTS}
    sum((weights[k] * oo[k]).sum() for k in au.OUTPUTS).backward()
    # kernels
    dev = torch.device("cuda:0")
    lk = {k: t.to(dev).requires_grad_(k != "opacities") for k, t in inp.items()}
    ok = _adapter(deg).forward(lk["extrinsics"], lk["intrinsics"], lk["coordinates"], lk["depths"], lk["opacities"],
                               lk["raw_gaussians"], (h, w), sh_rotation=dmat.float().to(dev))
    sum((weights[k].float().to(dev) * getattr(ok, k)).sum() for k in au.OUTPUTS).backward()
    for k in au.OUTPUTS:
        assert au.rel_err(getattr(ok, k).detach().cpu(), oo[k].detach()) < 1e-5, k
    for k in au.GRADS:
        assert au.rel_err(lk[k].grad.cpu(), lo[k].grad) < 1e-3, (k, au.rel_err(lk[k].grad.cpu(), lo[k].grad))


def test_adapter_output_feeds_the_decoder_without_a_relayout_copy():
    """End to end: adapter -> decoder_forward; the harmonics arrive in the rasterizer's layout (no .contiguous() copy)
    and gradients reach the raw Gaussians and the depths."""
    sys.path.insert(0, au.GOLDEN)
    import make_adapter_golden as mk
    from pf3plat_b200.render import decoder_forward
    dev = torch.device("cuda:0")
    b, v, h, w = 1, 2, 32, 32
    inp = {k: t.to(dev) for k, t in mk.make_inputs(11, b, v, h, w, 4).items()}
    inp["extrinsics"][..., :3, :3] = torch.eye(3, device=dev)   # cameras looking down +z, Gaussians 1..10 units ahead
    inp["extrinsics"][..., :3, 3] *= 0.1
    raw = inp["raw_gaussians"].clone().requires_grad_(True)
    dep = inp["depths"].clone().requires_grad_(True)
    g = _adapter(4).forward(inp["extrinsics"], inp["intrinsics"], inp["coordinates"], dep, inp["opacities"], raw, (h, w))
    flat = lambda t, n: t.reshape(b, -1, *t.shape[-n:]) if n else t.reshape(b, -1)
    harm = flat(g.harmonics, 2)
    assert harm.transpose(-1, -2).is_contiguous()
    ext = inp["extrinsics"].reshape(b, v, 4, 4)
    intr = inp["intrinsics"].reshape(b, v, 3, 3)
    near, far = torch.full((b, v), 0.5, device=dev), torch.full((b, v), 100.0, device=dev)
    color, _ = decoder_forward(flat(g.means, 1), flat(g.covariances, 2), harm, flat(g.opacities, 0), ext, intr, near, far,
                               (h, w), torch.zeros(3, device=dev))
    assert color.shape == (b, v, 3, h, w) and torch.isfinite(color).all() and color.abs().sum() > 0
    color.sum().backward()
    assert torch.isfinite(raw.grad).all() and raw.grad.abs().sum() > 0 and dep.grad.abs().sum() > 0


def test_record_adapter_timing():
    """Not a pass/fail speed gate: times the fused adapter (forward + backward) beside the reference's op sequence
    (the oracle's torch code, run on the GPU in fp32) at PF3plat's size -- 2 context views x 256x256 Gaussians -- and
    writes gpurun_out/adapter_timing.json when that directory exists."""
    import json
    sys.path.insert(0, au.GOLDEN)
    import make_adapter_golden as mk
    from oracle import adapter_oracle
    from pf3plat_b200.sh_rotation import sh_rotation_blocks
    dev = torch.device("cuda:0")
    b, v, h, w, deg = 1, 2, 256, 256, 4
    inp = {k: t.to(dev) for k, t in mk.make_inputs(3, b, v, h, w, deg).items()}
    dmat = sh_rotation_blocks(inp["extrinsics"][..., :3, :3].reshape(b, v, 3, 3), 25)
    adapter = _adapter(deg)

    def fused():
        raw = inp["raw_gaussians"].clone().requires_grad_(True)
        dep = inp["depths"].clone().requires_grad_(True)
        g = adapter.forward(inp["extrinsics"], inp["intrinsics"], inp["coordinates"], dep, inp["opacities"], raw, (h, w),
                            sh_rotation=dmat)
        (g.means.sum() + g.covariances.sum() + g.harmonics.sum()).backward()

    def unfused():
        raw = inp["raw_gaussians"].clone().requires_grad_(True)
        dep = inp["depths"].clone().requires_grad_(True)
        o = adapter_oracle.adapter_forward(inp["extrinsics"], inp["intrinsics"], inp["coordinates"], dep, inp["opacities"],
                                           raw, (h, w), deg, 0.5, 15.0, sh_rotation=dmat[:, :, None, None, None])
        (o["means"].sum() + o["covariances"].sum() + o["harmonics"].sum()).backward()

    res = {}
    for name, fn in (("fused_ms", fused), ("torch_op_sequence_ms", unfused)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
    res["gaussians"] = b * v * h * w
    print(res)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "adapter_timing.json"), "w") as f:
            json.dump(res, f)
    assert res["fused_ms"] > 0
