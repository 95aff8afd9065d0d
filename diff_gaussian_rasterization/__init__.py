"""Drop-in module for PF3plat's `from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer` (/root/reference/src/model/decoder/cuda_splatting.py:5-8).

Put the repository root on PYTHONPATH (or install it) and the reference's `render_cuda`,
`render_cuda_orthographic` and `render_depth_cuda` run unmodified on the sm_100a kernels of
pf3plat_b200/csrc.  See INTEGRATION.md.
"""
from pf3plat_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
