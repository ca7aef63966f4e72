"""gms_b200 -- host side (Python/PyTorch plumbing) of the B200-native mesh-Gaussian rasterizer.
The compute lives in libgms_b200.so (csrc/, hand-written CUDA for sm_100a) behind the C ABI of include/gms_b200.h."""
from . import _lib  # noqa: F401

__version__ = "0.1"
