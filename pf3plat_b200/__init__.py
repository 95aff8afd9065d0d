"""pf3plat_b200 -- B200-native (sm_100a) differentiable 3D-Gaussian-splatting rasterizer for the PF3plat
decoder path.  `import pf3plat_b200` does not load the CUDA library; the operators do, and fail loudly if
it has not been built (there is no CPU fallback)."""

__version__ = "0.1.0"
