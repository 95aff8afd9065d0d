"""Image metrics of the evaluation path.

`compute_psnr` and `compute_ssim` have the signatures and semantics of the reference's
(/root/reference/src/evaluation/metrics.py:11-19, :38-54) and run on the device (gs_metrics.cu); the reference's
SSIM loops over images on the CPU through scikit-image.  (`compute_lpips` needs the pretrained VGG weights of the
`lpips` package -- no network here -- and is out of scope.)"""
from __future__ import annotations

import torch

from . import _capi


@torch.no_grad()
def compute_psnr(ground_truth: torch.Tensor, predicted: torch.Tensor) -> torch.Tensor:
    """(batch, channel, height, width) x 2 -> (batch,) PSNR in dB; inputs are clipped to [0, 1]."""
    if ground_truth.shape != predicted.shape or ground_truth.dim() < 2:
        raise ValueError("compute_psnr expects two tensors of the same (batch, ...) shape")
    if predicted.device.type != "cuda":
        raise RuntimeError("pf3plat_b200.metrics needs CUDA tensors (there is no CPU fallback)")
    gt = ground_truth.detach().to(predicted.device, torch.float32).contiguous()
    pr = predicted.detach().to(torch.float32).contiguous()
    b = pr.shape[0]
    n = pr[0].numel() if b else 1
    L = _capi.lib()
    out = torch.empty((b,), dtype=torch.float32, device=pr.device)
    with torch.cuda.device(pr.device):
        scratch = torch.empty((max(1, int(L.gs_psnr_scratch_floats(b, n))),), dtype=torch.float32, device=pr.device)
        _capi.check(L.gs_psnr(gt.data_ptr(), pr.data_ptr(), b, n, scratch.data_ptr(), out.data_ptr(),
                              torch.cuda.current_stream(pr.device).cuda_stream))
    return out


@torch.no_grad()
def compute_ssim(ground_truth: torch.Tensor, predicted: torch.Tensor) -> torch.Tensor:
    """(batch, channel, height, width) x 2 -> (batch,) mean SSIM (Gaussian window, sigma 1.5, 11 taps, data range 1)."""
    if ground_truth.shape != predicted.shape or ground_truth.dim() != 4:
        raise ValueError("compute_ssim expects two (batch, channel, height, width) tensors of the same shape")
    if predicted.device.type != "cuda":
        raise RuntimeError("pf3plat_b200.metrics needs CUDA tensors (there is no CPU fallback)")
    gt = ground_truth.detach().to(predicted.device, torch.float32).contiguous()
    pr = predicted.detach().to(torch.float32).contiguous()
    b, c, h, w = pr.shape
    L = _capi.lib()
    out = torch.empty((b,), dtype=torch.float32, device=pr.device)
    with torch.cuda.device(pr.device):
        n = max(2, int(L.gs_ssim_scratch_floats(b, c, h, w)))
        scratch = torch.empty(((n + 1) // 2,), dtype=torch.float64, device=pr.device)
        _capi.check(L.gs_ssim(gt.data_ptr(), pr.data_ptr(), b, c, h, w, scratch.data_ptr(), out.data_ptr(),
                              torch.cuda.current_stream(pr.device).cuda_stream))
    return out.to(predicted.dtype)
