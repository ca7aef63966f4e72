"""Drop-in module name for the reference's import
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
(renderer/gaussian_renderer/__init__.py:14 and the three other renderers, line 16).
Put `gaussian-mesh-splatting_b200/` on sys.path (or `pip install -e` it) and the reference's renderer/ and
games/ packages run on the B200-native library unchanged."""
from gms_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,  # noqa: F401
                                 _RasterizeGaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
