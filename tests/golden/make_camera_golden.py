"""Generates tests/golden/camera_glue.npz by running the REFERENCE's own render_cuda
(/root/reference/src/model/decoder/cuda_splatting.py:47-127, imported unmodified from the read-only tree) against a
RECORDING stand-in for `diff_gaussian_rasterization`: the stub rasterizer stores the GaussianRasterizationSettings it is
handed for every view (viewmatrix, projmatrix, campos, tanfovx, tanfovy -- everything lines 64-112 compute) and returns
blank images.  The fixture therefore pins pf3plat_b200.cameras.make_view_batch / gs_view_batch to the reference's own glue.
Only runs in the build container (needs /root/reference).   Usage: python tests/golden/make_camera_golden.py"""
import importlib.util
import os
import sys
import types
from typing import NamedTuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
RECORDED: list = []


class _Settings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _Recorder:
    def __init__(self, raster_settings):
        self.s = raster_settings

    def __call__(self, means3D, means2D, shs=None, colors_precomp=None, opacities=None, cov3D_precomp=None, **kw):
        s = self.s
        RECORDED.append(dict(viewmatrix=s.viewmatrix.detach().clone(), projmatrix=s.projmatrix.detach().clone(),
                             campos=s.campos.detach().clone(), tanfov=torch.tensor([float(s.tanfovx), float(s.tanfovy)]),
                             means=means3D.detach().clone(), cov6=cov3D_precomp.detach().clone(),
                             bg=s.bg.detach().clone().float(), opacities=opacities.detach().clone(),
                             colors=(torch.zeros(0) if colors_precomp is None else colors_precomp.detach().clone()),
                             shs=(torch.zeros(0) if shs is None else shs.detach().clone()),
                             ints=torch.tensor([s.image_height, s.image_width, s.sh_degree, int(s.prefiltered), int(s.debug)]),
                             scale_modifier=torch.tensor(float(s.scale_modifier)),
                             means2D_is_zero_leaf=torch.tensor(int(bool((means2D == 0).all()) and means2D.requires_grad))))
        return torch.zeros(3, s.image_height, s.image_width), torch.zeros(means3D.shape[0], dtype=torch.int32)


def load_reference_render_cuda():
    stub = types.ModuleType("diff_gaussian_rasterization")
    stub.GaussianRasterizationSettings = _Settings
    stub.GaussianRasterizer = _Recorder
    saved = sys.modules.get("diff_gaussian_rasterization")
    sys.modules["diff_gaussian_rasterization"] = stub
    try:
        for name in ("src", "src.model", "src.model.decoder", "src.model.encoder", "src.model.encoder.costvolume", "src.geometry"):
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
            sys.modules[name] = m
        path = os.path.join(REF_ROOT, "src/model/decoder/cuda_splatting.py")
        spec = importlib.util.spec_from_file_location("src.model.decoder.cuda_splatting", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        if saved is not None:
            sys.modules["diff_gaussian_rasterization"] = saved
        else:
            del sys.modules["diff_gaussian_rasterization"]


def load_reference_decoder():
    """The reference's DecoderSplattingCUDA (/root/reference/src/model/decoder/decoder_splatting_cuda.py), on top of
    load_reference_render_cuda(); `src.dataset` (which pulls in every dataset class) is replaced by an empty stand-in --
    the decoder only reads dataset_cfg.background_color."""
    load_reference_render_cuda()
    ds = types.ModuleType("src.dataset")
    ds.DatasetCfg = type("DatasetCfg", (), {})
    sys.modules["src.dataset"] = ds
    for name, rel in (("src.model.types", "src/model/types.py"), ("src.model.decoder.decoder", "src/model/decoder/decoder.py"),
                      ("src.model.decoder.decoder_splatting_cuda", "src/model/decoder/decoder_splatting_cuda.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["src.model.decoder.decoder_splatting_cuda"], sys.modules["src.model.types"]


DECODER_KEYS = ("viewmatrix", "projmatrix", "campos", "tanfov", "means", "cov6", "bg", "opacities", "colors", "shs", "ints")


def run_reference_decoder(b=2, v=3, G=2, depth_mode="depth"):
    """Inputs and the recorded operator calls of one DecoderSplattingCUDA.forward (colour pass, then depth pass)."""
    dec_mod, types_mod = load_reference_decoder()
    ext, intr, near, far = make_cameras(21, b * v)
    g = torch.Generator().manual_seed(22)
    means = torch.randn(b, G, 3, generator=g).abs() + 0.5
    a = torch.randn(b, G, 3, 3, generator=g)
    cov = a @ a.transpose(-1, -2)
    sh = torch.randn(b, G, 3, 25, generator=g)
    opac = torch.rand(b, G, generator=g)
    cfg = types.SimpleNamespace(background_color=[0.1, 0.2, 0.3])
    dec = dec_mod.DecoderSplattingCUDA(dec_mod.DecoderSplattingCUDACfg(name="splatting_cuda"), cfg)
    RECORDED.clear()
    r4 = lambda t: t.reshape(b, v, *t.shape[1:])
    dec.forward(types_mod.Gaussians(means, cov, sh, opac), r4(ext), r4(intr), r4(near), r4(far), (16, 24), depth_mode=depth_mode)
    rec = {k: [r[k] for r in RECORDED] for k in DECODER_KEYS}
    RECORDED.clear()
    inputs = dict(extrinsics=r4(ext), intrinsics=r4(intr), near=r4(near), far=r4(far), means=means, covariances=cov, sh=sh,
                  opacities=opac)
    return inputs, rec


def ortho_inputs():
    ext = make_cameras(31, 1)[0]
    return dict(extrinsics=ext, width=torch.tensor([2.5]), height=torch.tensor([1.75]), near=torch.tensor([0.0]),
                far=torch.tensor([7.0]), bg=torch.tensor([[0.3, 0.1, 0.2]]))


def make_cameras(seed, B):
    g = torch.Generator().manual_seed(seed)
    ext = torch.eye(4).repeat(B, 1, 1)
    ext[:, :3, :3] = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    ext[:, :3, 3] = torch.randn(B, 3, generator=g)
    intr = torch.eye(3).repeat(B, 1, 1)
    intr[:, 0, 0] = 0.6 + 0.6 * torch.rand(B, generator=g)
    intr[:, 1, 1] = 0.6 + 0.6 * torch.rand(B, generator=g)
    intr[:, 0, 2] = 0.5 + 0.03 * torch.randn(B, generator=g)
    intr[:, 1, 2] = 0.5 + 0.03 * torch.randn(B, generator=g)
    near = 0.3 + torch.rand(B, generator=g)
    far = 40 + 100 * torch.rand(B, generator=g)
    return ext, intr, near, far


def main():
    mod = load_reference_render_cuda()
    B, G = 9, 2
    ext, intr, near, far = make_cameras(12, B)
    g = torch.Generator().manual_seed(13)
    means = torch.randn(B, G, 3, generator=g)
    a = torch.randn(B, G, 3, 3, generator=g)
    cov = a @ a.transpose(-1, -2)
    sh = torch.randn(B, G, 3, 25, generator=g)
    opac = torch.rand(B, G, generator=g)
    arrays = dict(extrinsics=ext.numpy(), intrinsics=intr.numpy(), near=near.numpy(), far=far.numpy(),
                  means=means.numpy(), covariances=cov.numpy())
    for tag, si in (("si", True), ("raw", False)):
        RECORDED.clear()
        mod.render_cuda(ext, intr, near, far, (16, 24), torch.zeros(B, 3), means, cov, sh, opac, scale_invariant=si)
        assert len(RECORDED) == B
        for k in ("viewmatrix", "projmatrix", "campos", "tanfov", "means", "cov6"):
            arrays[f"{tag}_{k}"] = torch.stack([r[k] for r in RECORDED]).numpy()
    # the depth renders: what render_depth_cuda (:226-269) hands the op in each of its modes (fake colours, bg, flags)
    for mode in ("depth", "disparity", "relative_disparity", "log"):
        RECORDED.clear()
        mod.render_depth_cuda(ext, intr, near, far, (16, 24), means.abs() + 0.5, cov, opac, mode=mode)
        assert len(RECORDED) == B
        for k in ("colors", "bg", "opacities", "ints", "scale_modifier", "means2D_is_zero_leaf"):
            arrays[f"depth_{mode}_{k}"] = torch.stack([r[k] for r in RECORDED]).numpy()
    # and the colour render's remaining arguments
    RECORDED.clear()
    mod.render_cuda(ext, intr, near, far, (16, 24), torch.rand(B, 3, generator=torch.Generator().manual_seed(14)), means, cov, sh, opac)
    for k in ("shs", "bg", "opacities", "ints", "scale_modifier", "means2D_is_zero_leaf"):
        arrays[f"color_{k}"] = torch.stack([r[k] for r in RECORDED]).numpy()
    arrays["sh"] = sh.numpy()
    arrays["opacities"] = opac.numpy()
    # the fake-orthographic render (render_cuda_orthographic, :130-220; batch of one, as its 0-dim/1-element tensor
    # tanfov arguments require): everything it hands the op
    for k, t in ortho_inputs().items():
        arrays[f"ortho_in_{k}"] = t.numpy()
    RECORDED.clear()
    oi = ortho_inputs()
    mod.render_cuda_orthographic(oi["extrinsics"], oi["width"], oi["height"], oi["near"], oi["far"], (16, 24), oi["bg"],
                                 means[:1], cov[:1], sh[:1], opac[:1])
    assert len(RECORDED) == 1
    for k in ("viewmatrix", "projmatrix", "campos", "tanfov", "means", "cov6", "shs", "bg", "opacities", "ints",
              "scale_modifier", "means2D_is_zero_leaf"):
        arrays[f"ortho_{k}"] = torch.stack([r[k] for r in RECORDED]).numpy()
    # the decoder on top (decoder_splatting_cuda.py:35-91): 2 scenes x 3 views, colour pass then depth pass
    inputs, rec = run_reference_decoder()
    for k, t in inputs.items():
        arrays[f"dec_in_{k}"] = t.numpy()
    for k, lst in rec.items():
        for half, sl in (("color", slice(0, 6)), ("depth", slice(6, 12))):
            arrays[f"dec_{half}_{k}"] = torch.stack(lst[sl]).numpy()
    path = os.path.join(HERE, "camera_glue.npz")
    np.savez_compressed(path, **arrays)
    print({k: v.shape for k, v in arrays.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
