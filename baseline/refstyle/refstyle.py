"""Python front end of the ref-style comparator (baseline/refstyle/refstyle.cu) -- NOT the product.

Exposes the same `GaussianRasterizationSettings` / `GaussianRasterizer` surface as the stock extension so that the
reference's unmodified `render()` can run on it; bench.py and tests/test_gpu_refstyle.py use it to put a number and a
parity check next to the product.  It is a labelled stand-in for the stock rasterizer (whose source is an empty
submodule of the reference checkout), built from SURVEY.md Appendix A with the stock work decomposition.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_PKG = os.path.join(_ROOT, "gaussian-mesh-splatting_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from gms_b200 import _lib as _abi                      # ctypes struct definitions of include/gms_b200.h only
from gms_b200.rasterizer import GaussianRasterizationSettings, _dev_f32, _opt, _ptr   # the settings tuple and tensor helpers

LIB_PATH = os.path.join(_HERE, "librefstyle.so")
SRC = os.path.join(_HERE, "refstyle.cu")
_L = None


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(_ROOT, "include", "gms_b200.h"), os.path.join(_PKG, "csrc", "gms_preprocess.cuh"),
            os.path.join(_PKG, "csrc", "gms_common.cuh")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # stock build flags: -O3, no --use_fast_math
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
           "-ccbin", "/usr/bin/g++", "-o", LIB_PATH, SRC]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    global _L
    if _L is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built (python -c 'import __graft_entry__ as g; g.build()')")
        L = C.CDLL(LIB_PATH)
        L.refstyle_last_error.restype = C.c_char_p
        L.refstyle_backward_scratch_bytes.restype = C.c_size_t
        L.refstyle_backward_scratch_bytes.argtypes = [C.c_int32]
        L.refstyle_rasterize_forward.argtypes = [C.POINTER(_abi.RasterSettings), C.POINTER(_abi.RasterInputs), C.POINTER(_abi.RasterOutputs),
                                                 _abi.ALLOC_FN, C.c_void_p, C.POINTER(_abi.RasterSaved), C.c_void_p]
        L.refstyle_rasterize_backward.argtypes = [C.POINTER(_abi.RasterSettings), C.POINTER(_abi.RasterInputs), C.c_void_p,
                                                  C.POINTER(_abi.RasterSaved), C.c_void_p, C.c_void_p, C.POINTER(_abi.RasterGrads),
                                                  C.c_void_p, C.c_void_p]
        _L = L
    return _L


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().refstyle_last_error().decode()}")


def _settings(rs, device, keep):
    s = _abi.RasterSettings()
    s.image_height, s.image_width = int(rs.image_height), int(rs.image_width)
    s.tanfovx, s.tanfovy, s.scale_modifier, s.sh_degree = float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier), int(rs.sh_degree)
    s.prefiltered, s.debug, s.antialiasing = int(bool(rs.prefiltered)), int(bool(rs.debug)), int(bool(rs.antialiasing))
    t = [_dev_f32(x, device) for x in (rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos)]
    keep += t
    s.bg, s.viewmatrix, s.projmatrix, s.campos = (x.data_ptr() for x in t)
    return s


class _RefStyleRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        L = lib()
        dev = means3D.device
        P, H, W = means3D.shape[0], int(rs.image_height), int(rs.image_width)
        f = lambda t: None if _opt(t) is None else _dev_f32(t.detach(), dev)
        m3, op, shs, col, sc, rot, cov = f(means3D), f(opacities), f(sh), f(colors_precomp), f(scales), f(rotations), f(cov3Ds_precomp)
        M = shs.shape[1] if shs is not None else 0
        keep = []
        s = _settings(rs, dev, keep)
        i = _abi.RasterInputs()
        i.P, i.M, i.means3D, i.opacities = P, M, _ptr(m3), _ptr(op)
        i.shs, i.colors_precomp, i.scales, i.rotations, i.cov3D_precomp = _ptr(shs), _ptr(col), _ptr(sc), _ptr(rot), _ptr(cov)
        # stock wrapper: torch::full / torch::zeros outputs, three resizable byte tensors served through a callback
        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        invdepth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
        bufs = {}

        def _alloc(user, which, nbytes):
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            bufs[int(which)] = t
            return t.data_ptr()

        cb = _abi.ALLOC_FN(_alloc)
        o = _abi.RasterOutputs(color.data_ptr(), radii.data_ptr(), invdepth.data_ptr(), 0)
        saved = _abi.RasterSaved()
        with torch.cuda.device(dev):
            _check(L.refstyle_rasterize_forward(C.byref(s), C.byref(i), C.byref(o), cb, None, C.byref(saved),
                                                torch.cuda.current_stream(dev).cuda_stream), "refstyle_rasterize_forward")
        del cb
        ctx.rs, ctx.bufs, ctx.N, ctx.dims, ctx.keep = rs, bufs, int(saved.num_rendered), (P, M), keep
        ctx.save_for_backward(m3, op, shs, col, sc, rot, cov, radii)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)
        return color, radii, invdepth

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth):
        L = lib()
        m3, op, shs, col, sc, rot, cov, radii = ctx.saved_tensors
        dev = m3.device
        P, M = ctx.dims
        rs = ctx.rs
        keep = []
        s = _settings(rs, dev, keep)
        i = _abi.RasterInputs()
        i.P, i.M, i.means3D, i.opacities = P, M, _ptr(m3), _ptr(op)
        i.shs, i.colors_precomp, i.scales, i.rotations, i.cov3D_precomp = _ptr(shs), _ptr(col), _ptr(sc), _ptr(rot), _ptr(cov)
        saved = _abi.RasterSaved()
        b = ctx.bufs
        saved.geom, saved.binning, saved.image = _ptr(b.get(0)), _ptr(b.get(1)), _ptr(b.get(2))
        saved.num_rendered = ctx.N
        if g_color is None:
            g_color = torch.zeros((3, int(rs.image_height), int(rs.image_width)), device=dev)
        gc = _dev_f32(g_color, dev)
        gd = None if g_depth is None else _dev_f32(g_depth, dev)
        e = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
        g_m3, g_m2, g_op = e(P, 3), e(P, 3), e(P, 1)
        g_sh = e(P, M, 3) if shs is not None else None
        g_col = e(P, 3) if col is not None else None
        g_sc = e(P, 3) if sc is not None else None
        g_rot = e(P, 4) if rot is not None else None
        g_cov = e(P, 6) if cov is not None else None
        gr = _abi.RasterGrads(_ptr(g_m3), _ptr(g_m2), _ptr(g_op), _ptr(g_sh), _ptr(g_col), _ptr(g_sc), _ptr(g_rot), _ptr(g_cov))
        scratch = torch.empty(int(L.refstyle_backward_scratch_bytes(P)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _check(L.refstyle_rasterize_backward(C.byref(s), C.byref(i), radii.data_ptr(), C.byref(saved), gc.data_ptr(), _ptr(gd),
                                                 C.byref(gr), scratch.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                   "refstyle_rasterize_backward")
        return g_m3, g_m2, g_sh, g_col, g_op, g_sc, g_rot, g_cov, None


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return _RefStyleRasterize.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                        self.raster_settings)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
