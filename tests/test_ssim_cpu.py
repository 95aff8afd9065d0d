"""Closed-form checks of the SSIM oracle (oracle/ssim_oracle.py); scikit-image itself is absent from this image."""
import numpy as np
import pytest

from oracle import ssim_oracle as so


def _weights():
    k = np.arange(-5, 6, dtype=np.float64)
    w = np.exp(-0.5 * k * k / 1.5 ** 2)
    return w / w.sum()


def test_identical_images_give_one():
    rng = np.random.default_rng(0)
    x = rng.random((2, 3, 40, 33)).astype(np.float32)
    np.testing.assert_allclose(so.compute_ssim(x, x), 1.0, atol=1e-12)


def test_constant_images_reduce_to_the_luminance_term():
    a, b = 0.3, 0.7
    x, y = np.full((1, 1, 20, 20), a), np.full((1, 1, 20, 20), b)
    c1 = 0.01 ** 2
    np.testing.assert_allclose(so.compute_ssim(x, y), (2 * a * b + c1) / (a * a + b * b + c1), rtol=1e-12)


def test_single_retained_pixel_matches_brute_force_window_sums():
    """An 11x11 image keeps exactly one SSIM value (the centre), whose window is the whole image."""
    rng = np.random.default_rng(1)
    x, y = rng.random((11, 11)), rng.random((11, 11))
    w2 = np.outer(_weights(), _weights())
    m = lambda a: float((w2 * a).sum())
    ux, uy = m(x), m(y)
    n = 121 / 120
    vx, vy, vxy = n * (m(x * x) - ux * ux), n * (m(y * y) - uy * uy), n * (m(x * y) - ux * uy)
    c1, c2 = 1e-4, 9e-4
    want = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    np.testing.assert_allclose(so.ssim_channel(x, y), want, rtol=1e-12)


def test_channel_axis_means_average_of_channel_means_and_too_small_images_raise():
    rng = np.random.default_rng(2)
    x, y = rng.random((1, 3, 16, 21)), rng.random((1, 3, 16, 21))
    per = [so.ssim_channel(x[0, c], y[0, c]) for c in range(3)]
    np.testing.assert_allclose(so.compute_ssim(x, y)[0], np.mean(per), rtol=1e-14)
    with pytest.raises(ValueError):
        so.compute_ssim(np.zeros((1, 1, 10, 30)), np.zeros((1, 1, 10, 30)))
