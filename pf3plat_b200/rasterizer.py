"""Host-side mirror of the reference's rasterizer operator, backed by the C ABI (include/gsplat_b200.h).

Surface replaced: the `diff_gaussian_rasterization` module PF3plat imports
(/root/reference/src/model/decoder/cuda_splatting.py:5-8) -- `GaussianRasterizationSettings` (12-field
NamedTuple, constructed at :99-112 and :192-205) and `GaussianRasterizer` (nn.Module called with keywords at
:117-124 and :210-217, returning the 2-tuple `(color, radii)`).  Same names, argument meaning and error
behaviour; PyTorch is used for device memory, streams and autograd plumbing only.

`rasterize_batch` is the batched form of the same operator (V views of S scenes per call) used by
`pf3plat_b200.render`; the per-view reference call is its S = V = 1 case.
"""
from __future__ import annotations

import atexit
import ctypes
import threading
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _capi
from ._capi import GsConfig, GsInGrads, GsInputs, GsOutGrads, GsOutputs, GsStats


class GaussianRasterizationSettings(NamedTuple):
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


# ---------------------------------------------------------------------------------------------------------
# contexts: one GsContext per (device, stream)
# ---------------------------------------------------------------------------------------------------------
_contexts: dict[tuple[int, int], ctypes.c_void_p] = {}
_locks: dict[tuple[int, int], threading.RLock] = {}
_registry_lock = threading.Lock()


def _key(device: torch.device, stream_ptr: int) -> tuple[int, int]:
    return (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr)


def _context(device: torch.device, stream_ptr: int) -> ctypes.c_void_p:
    key = _key(device, stream_ptr)
    with _registry_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = ctypes.c_void_p()
            _capi.check(_capi.lib().gs_context_create(ctypes.byref(ctx)))
            _contexts[key] = ctx
            _locks[key] = threading.RLock()
    return ctx


def _context_lock(device: torch.device, stream_ptr: int) -> threading.RLock:
    """A GsContext serves one call at a time (include/gsplat_b200.h): the forward of one thread and the backward the
    autograd engine runs on another take this lock around their gs_* calls."""
    _context(device, stream_ptr)
    return _locks[_key(device, stream_ptr)]


def trim_memory(device=None) -> None:
    """Hands the cached blocks of every context's private memory pool on `device` (all devices if None) back to the
    driver -- the counterpart of torch.cuda.empty_cache() for the memory this library holds outside torch's allocator."""
    with _registry_lock:
        items = list(_contexts.items())
    for (dev_index, _), ctx in items:
        if device is None or torch.device(device).index in (None, dev_index):
            with torch.cuda.device(dev_index):
                _capi.check(_capi.lib().gs_context_trim(ctx))


def destroy_contexts() -> None:
    """Destroys every cached GsContext (registered with atexit; also usable from tests)."""
    with _registry_lock:
        items = list(_contexts.items())
        _contexts.clear()
        _locks.clear()
    for (dev_index, _), ctx in items:
        try:
            with torch.cuda.device(dev_index):
                _capi.lib().gs_context_destroy(ctx)
        except Exception:
            pass


atexit.register(destroy_contexts)


def current_context(device=None) -> ctypes.c_void_p:
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(device):
        return _context(device, torch.cuda.current_stream(device).cuda_stream)


def last_stats(device=None) -> dict:
    st = GsStats()
    _capi.check(_capi.lib().gs_get_stats(current_context(device), ctypes.byref(st)))
    return {k: getattr(st, k) for k, _ in GsStats._fields_ if not k.endswith("_")}


def set_profiling(enabled: bool, device=None) -> None:
    _capi.check(_capi.lib().gs_set_profiling(current_context(device), int(enabled)))


def stage_ms(device=None) -> dict:
    arr = (ctypes.c_float * _capi.GS_NUM_STAGES)()
    _capi.check(_capi.lib().gs_get_stage_ms(current_context(device), arr))
    return dict(zip(_capi.STAGE_NAMES, [float(x) for x in arr]))


class _SavedHandle:
    """Owns a GsSaved*; released on the stream of the forward when the autograd graph lets go of it."""

    def __init__(self, ctx, ptr, stream_ptr, device):
        self.ctx, self.ptr, self.stream_ptr, self.device = ctx, ptr, stream_ptr, device

    def __del__(self):
        try:
            if self.ptr:
                with torch.cuda.device(self.device):
                    _capi.lib().gs_saved_free(self.ctx, self.ptr, self.stream_ptr)
                self.ptr = None
        except Exception:
            pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: Optional[torch.Tensor], device=None) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if device is not None and t.device != device:
        t = t.to(device)
    return t.detach().to(torch.float32).contiguous()


class BatchSettings(NamedTuple):
    """GaussianRasterizationSettings for V views at once (see include/gsplat_b200.h: GsConfig)."""
    image_height: int
    image_width: int
    viewmatrix: torch.Tensor                 # (V,4,4) transposed
    projmatrix: torch.Tensor                 # (V,4,4) transposed
    campos: torch.Tensor                     # (V,3)
    bg: torch.Tensor                         # (V,3)
    sh_degree: int
    tanfov: Optional[torch.Tensor] = None    # (V,2) device tensor, or
    tanfovx: float = 0.0                     # one host pair for all views
    tanfovy: float = 0.0
    view_scale: Optional[torch.Tensor] = None  # (V,)
    scale_modifier: float = 1.0
    prefiltered: bool = False
    with_depth: bool = False
    debug: bool = False
    tuning: int = 0                          # GS_TUNE_* knobs (testing)


def _make_config(bs: BatchSettings, S: int, P: int, M: int, keep: list) -> GsConfig:
    V = bs.viewmatrix.shape[0]
    dev = bs.viewmatrix.device
    cfg = GsConfig()
    cfg.P, cfg.S, cfg.V, cfg.M = P, S, V, M
    cfg.sh_degree = int(bs.sh_degree)
    cfg.image_height, cfg.image_width = int(bs.image_height), int(bs.image_width)
    cfg.flags = (_capi.GS_FLAG_DEPTH if bs.with_depth else 0) | (_capi.GS_FLAG_PREFILTERED if bs.prefiltered else 0)
    cfg.tanfovx, cfg.tanfovy = float(bs.tanfovx), float(bs.tanfovy)
    cfg.scale_modifier = float(bs.scale_modifier)
    cfg.tuning = int(bs.tuning)
    tensors = {
        "viewmatrix": _f32c(bs.viewmatrix).reshape(V, 16), "projmatrix": _f32c(bs.projmatrix, dev).reshape(V, 16),
        "campos": _f32c(bs.campos, dev).reshape(V, 3), "bg": _f32c(bs.bg, dev).reshape(V, 3),
        "tanfov": None if bs.tanfov is None else _f32c(bs.tanfov, dev).reshape(V, 2),
        "view_scale": None if bs.view_scale is None else _f32c(bs.view_scale, dev).reshape(V),
    }
    for k, t in tensors.items():
        setattr(cfg, k, _ptr(t))
        keep.append(t)
    return cfg


class _RasterizeBatch(torch.autograd.Function):
    """The reference's _RasterizeGaussians, batched over views."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, bs: BatchSettings):
        dev = means3D.device
        if dev.type != "cuda":
            raise RuntimeError("pf3plat_b200 rasterizer needs CUDA tensors (there is no CPU fallback)")
        S, P = means3D.shape[0], means3D.shape[1]
        V = bs.viewmatrix.shape[0]
        M = 0 if shs is None else shs.shape[2]
        H, W = int(bs.image_height), int(bs.image_width)
        keep: list = []
        with torch.cuda.device(dev):
            stream_ptr = torch.cuda.current_stream(dev).cuda_stream
            gctx = _context(dev, stream_ptr)
            cfg = _make_config(bs, S, P, M, keep)
            ins = {
                "means3D": _f32c(means3D), "opacities": _f32c(opacities), "shs": _f32c(shs),
                "colors_precomp": _f32c(colors_precomp), "scales": _f32c(scales), "rotations": _f32c(rotations),
                "cov3D_precomp": _f32c(cov3D_precomp),
            }
            gin = GsInputs(**{k: _ptr(t) for k, t in ins.items()})
            color = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((V, P), dtype=torch.int32, device=dev)
            depth = torch.empty((V, H, W), dtype=torch.float32, device=dev) if bs.with_depth else None
            gout = GsOutputs(_ptr(color), _ptr(radii), _ptr(depth))
            needs_grad = any(t is not None and t.requires_grad for t in
                             (means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp))
            saved = ctypes.c_void_p()
            with _context_lock(dev, stream_ptr):
                rc = _capi.lib().gs_forward(gctx, ctypes.byref(cfg), ctypes.byref(gin), ctypes.byref(gout),
                                            ctypes.byref(saved) if needs_grad else None, stream_ptr)
                _capi.check(rc)
        if needs_grad:
            ctx.handle = _SavedHandle(gctx, saved, stream_ptr, dev)
            ctx.bs = bs
            ctx.ins = ins            # contiguous fp32 inputs the backward kernels re-read
            # `ins` aliases the caller's tensors when they already are fp32-contiguous; autograd's own version check is
            # bypassed for them, so do it by hand: an in-place update between forward and backward must raise, as it
            # does with the reference op (which hands its inputs to save_for_backward)
            srcs = {"means3D": means3D, "opacities": opacities, "shs": shs, "colors_precomp": colors_precomp,
                    "scales": scales, "rotations": rotations, "cov3D_precomp": cov3D_precomp}
            ctx.versions = [(k, t, t._version) for k, t in srcs.items() if t is not None]
            ctx.cfg_keep = keep
            ctx.dims = (S, P, V, M, H, W)
            ctx.want_means2D = means2D is not None and means2D.requires_grad
        ctx.mark_non_differentiable(radii)
        if bs.with_depth:
            return color, radii, depth
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, grad_radii=None, grad_depth=None):
        S, P, V, M, H, W = ctx.dims
        ins = ctx.ins
        for name, t, version in ctx.versions:
            if t._version != version:
                raise RuntimeError(f"one of the variables needed for gradient computation has been modified by an inplace "
                                   f"operation: {name} is at version {t._version}; expected version {version} instead")
        dev = ins["means3D"].device
        bs = ctx.bs
        keep: list = []
        with torch.cuda.device(dev):
            stream_ptr = torch.cuda.current_stream(dev).cuda_stream
            gctx = _context(dev, stream_ptr)
            cfg = _make_config(bs, S, P, M, keep)
            gin = GsInputs(**{k: _ptr(t) for k, t in ins.items()})
            gcol = _f32c(grad_color) if grad_color is not None else torch.zeros((V, 3, H, W), device=dev)
            gdep = _f32c(grad_depth) if (bs.with_depth and grad_depth is not None) else None
            og = GsOutGrads(_ptr(gcol), _ptr(gdep))
            e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            g = {
                "dL_dmeans3D": e(S, P, 3), "dL_dmeans2D": e(V, P, 3) if ctx.want_means2D else None,
                "dL_dshs": e(S, P, M, 3) if ins["shs"] is not None else None,
                "dL_dcolors": e(V, P, 3) if ins["colors_precomp"] is not None else None,
                "dL_dopacities": e(*ins["opacities"].shape),
                "dL_dscales": e(S, P, 3) if ins["scales"] is not None else None,
                "dL_drotations": e(S, P, 4) if ins["scales"] is not None else None,
                "dL_dcov3D": e(S, P, 6) if ins["cov3D_precomp"] is not None else None,
            }
            ig = GsInGrads(**{k: _ptr(t) for k, t in g.items()})
            with _context_lock(dev, stream_ptr):
                rc = _capi.lib().gs_backward(gctx, ctypes.byref(cfg), ctypes.byref(gin), ctx.handle.ptr, ctypes.byref(og),
                                             ctypes.byref(ig), stream_ptr)
                _capi.check(rc)
        return (g["dL_dmeans3D"], g["dL_dmeans2D"], g["dL_dshs"], g["dL_dcolors"], g["dL_dopacities"],
                g["dL_dscales"], g["dL_drotations"], g["dL_dcov3D"], None)


def rasterize_batch(bs: BatchSettings, means3D, opacities, shs=None, colors_precomp=None, scales=None,
                    rotations=None, cov3D_precomp=None, means2D=None):
    """Batched operator.  Shapes: means3D (S,P,3); opacities (S,P) or (S,P,1); shs (S,P,M,3); colors_precomp
    (V,P,3); scales (S,P,3); rotations (S,P,4); cov3D_precomp (S,P,6); means2D (V,P,3) (gradient sink only).
    Returns (color (V,3,H,W), radii (V,P)[, depth (V,H,W)])."""
    _check_exclusive(shs, colors_precomp, scales, rotations, cov3D_precomp)
    # means2D exists in the reference op only as a sink for the screen-space gradient (cuda_splatting.py:93-97); the
    # batched entry neither allocates nor fills it unless the caller passes one
    return _RasterizeBatch.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, bs)


def _check_exclusive(shs, colors_precomp, scales, rotations, cov3D_precomp):
    # same checks and messages as the reference op's GaussianRasterizer.forward
    if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")


class GaussianRasterizer(nn.Module):
    """Drop-in for diff_gaussian_rasterization.GaussianRasterizer (cuda_splatting.py:113-124, :206-217)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def _batch_settings(self, device, with_depth=False) -> BatchSettings:
        rs = self.raster_settings
        f = lambda x: float(x)  # accepts python floats and 0-dim tensors (cuda_splatting.py:102-103 vs :195-196)
        return BatchSettings(
            image_height=int(rs.image_height), image_width=int(rs.image_width),
            viewmatrix=rs.viewmatrix.reshape(1, 4, 4).to(device), projmatrix=rs.projmatrix.reshape(1, 4, 4),
            campos=rs.campos.reshape(1, 3), bg=rs.bg.reshape(1, 3), sh_degree=int(rs.sh_degree),
            tanfovx=f(rs.tanfovx), tanfovy=f(rs.tanfovy), scale_modifier=float(rs.scale_modifier),
            prefiltered=bool(rs.prefiltered), with_depth=with_depth, debug=bool(rs.debug))

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            dev = positions.device
            P = positions.shape[0]
            keep: list = []
            with torch.cuda.device(dev):
                stream_ptr = torch.cuda.current_stream(dev).cuda_stream
                gctx = _context(dev, stream_ptr)
                cfg = _make_config(self._batch_settings(dev), 1, P, 0, keep)
                pos = _f32c(positions)
                present = torch.empty((P,), dtype=torch.uint8, device=dev)
                with _context_lock(dev, stream_ptr):
                    _capi.check(_capi.lib().gs_mark_visible(gctx, ctypes.byref(cfg), pos.data_ptr(), present.data_ptr(),
                                                            stream_ptr))
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, return_depth: bool = False):
        """Returns (color (3,H,W), radii (P,)) like the reference op; `return_depth=True` (opt-in, not used by
        PF3plat) appends the composited camera-space depth (H,W)."""
        _check_exclusive(shs, colors_precomp, scales, rotations, cov3D_precomp)
        u = lambda t: None if t is None else t.unsqueeze(0)
        bs = self._batch_settings(means3D.device, with_depth=return_depth)
        out = _RasterizeBatch.apply(u(means3D), u(means2D), u(shs), u(colors_precomp), u(opacities), u(scales),
                                    u(rotations), u(cov3D_precomp), bs)
        if return_depth:
            return out[0][0], out[1][0], out[2][0]
        return out[0][0], out[1][0]
