"""CPU restatement of the reference's SSIM metric, for the parity tests of gs_ssim.

TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

/root/reference/src/evaluation/metrics.py:38-54 calls, per image,
    skimage.metrics.structural_similarity(gt, hat, win_size=11, gaussian_weights=True, channel_axis=0, data_range=1.0)
scikit-image is a third-party dependency (unpinned in /root/reference/requirements.txt) that is ABSENT from this image, so
this file restates its published algorithm (skimage/metrics/_structural_similarity.py, v0.19-0.24: unchanged in this
respect) on top of the very filter it uses, scipy.ndimage.gaussian_filter (scipy IS installed):
  * gaussian_weights=True: sigma = 1.5, truncate = 3.5 -> radius int(3.5*1.5+0.5) = 5, 11 taps, mode 'reflect';
  * use_sample_covariance=True (default, not overridden by the reference): cov_norm = NP/(NP-1), NP = win_size**2 = 121;
  * K1 = 0.01, K2 = 0.03, C1 = (K1*R)^2, C2 = (K2*R)^2 with R = data_range = 1;
  * S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), cropped by (win_size-1)//2 = 5 on each
    side, mean in float64; channel_axis: the mean of the per-channel means.
PARITY UNPINNED against scikit-image itself (absent); pinned by closed-form cases in tests/test_ssim_cpu.py
(identical images -> 1, constant images -> the luminance term, brute-force window sums at single pixels).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter

SIGMA, TRUNCATE, WIN = 1.5, 3.5, 11


def ssim_channel(x: np.ndarray, y: np.ndarray, dtype=np.float64) -> float:
    x, y = x.astype(dtype, copy=False), y.astype(dtype, copy=False)
    if min(x.shape) < WIN:
        raise ValueError("win_size exceeds image extent.")
    f = lambda a: gaussian_filter(a, sigma=SIGMA, truncate=TRUNCATE, mode="reflect")
    cov_norm = WIN * WIN / (WIN * WIN - 1.0)
    ux, uy = f(x), f(y)
    uxx, uyy, uxy = f(x * x), f(y * y), f(x * y)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    pad = (WIN - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean(dtype=np.float64))


def compute_ssim(ground_truth: np.ndarray, predicted: np.ndarray, dtype=np.float64) -> np.ndarray:
    """(batch, channel, h, w) x 2 -> (batch,)   (metrics.py:38-54)."""
    return np.array([np.mean([ssim_channel(g[c], p[c], dtype) for c in range(g.shape[0])])
                     for g, p in zip(ground_truth, predicted)], dtype=np.float64)
