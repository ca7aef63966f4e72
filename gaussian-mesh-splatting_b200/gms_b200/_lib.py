"""ctypes binding of libgms_b200.so (C ABI declared in include/gms_b200.h).

The product path REQUIRES the CUDA library: there is no CPU or PyTorch fallback.  Importing this module
is cheap; the first call that needs the library raises GmsLibraryError if it has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `python gaussian-mesh-splatting_b200/build.py`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgms_b200.so")

FORWARD_ONLY = 1        # gms_raster_outputs.flags: no backward will follow (skip the survivor lists)
GMS_OK, GMS_E_ARG, GMS_E_CUDA, GMS_E_ALLOC, GMS_E_UNSUPPORTED = 0, -1, -2, -3, -4
BUF_GEOM, BUF_BINNING, BUF_IMAGE = 0, 1, 2

c_float_p = C.c_void_p   # raw device pointers travel as integers


class GmsLibraryError(RuntimeError):
    pass


class RasterSettings(C.Structure):
    _fields_ = [("image_height", C.c_int32), ("image_width", C.c_int32), ("tanfovx", C.c_float),
                ("tanfovy", C.c_float), ("bg", C.c_void_p), ("scale_modifier", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("sh_degree", C.c_int32),
                ("campos", C.c_void_p), ("prefiltered", C.c_int32), ("debug", C.c_int32),
                ("antialiasing", C.c_int32)]


class RasterInputs(C.Structure):
    _fields_ = [("P", C.c_int32), ("M", C.c_int32), ("means3D", C.c_void_p), ("opacities", C.c_void_p),
                ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("scales", C.c_void_p),
                ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p)]


class RasterOutputs(C.Structure):
    _fields_ = [("out_color", C.c_void_p), ("radii", C.c_void_p), ("out_invdepth", C.c_void_p), ("flags", C.c_int32)]


class RasterSaved(C.Structure):
    _fields_ = [("geom", C.c_void_p), ("binning", C.c_void_p), ("image", C.c_void_p),
                ("num_rendered", C.c_int64), ("num_visible", C.c_int64), ("binning_capacity", C.c_int64),
                ("flags", C.c_int32)]


class RasterGrads(C.Structure):
    _fields_ = [("dL_dmeans3D", C.c_void_p), ("dL_dmeans2D", C.c_void_p), ("dL_dopacities", C.c_void_p),
                ("dL_dshs", C.c_void_p), ("dL_dcolors_precomp", C.c_void_p), ("dL_dscales", C.c_void_p),
                ("dL_drotations", C.c_void_p), ("dL_dcov3D_precomp", C.c_void_p), ("dL_dcolors_sh", C.c_void_p)]


class DebugViews(C.Structure):
    _fields_ = [("means2D", C.c_void_p), ("depths", C.c_void_p), ("cov3D", C.c_void_p),
                ("conic_opacity", C.c_void_p), ("rgb", C.c_void_p), ("clamped", C.c_void_p),
                ("tiles_touched", C.c_void_p), ("point_list", C.c_void_p), ("tile_keys", C.c_void_p),
                ("ranges", C.c_void_p), ("final_T", C.c_void_p), ("n_contrib", C.c_void_p), ("dgeom", C.c_void_p)]


class ExpandArgs(C.Structure):
    _fields_ = [("V", C.c_int32), ("F", C.c_int32), ("K", C.c_int32), ("vertices", C.c_void_p),
                ("faces", C.c_void_p), ("triangles_in", C.c_void_p), ("alpha_raw", C.c_void_p),
                ("scale_raw", C.c_void_p), ("eps", C.c_float), ("alpha", C.c_void_p),
                ("triangles", C.c_void_p), ("xyz", C.c_void_p), ("scaling_log", C.c_void_p),
                ("rotation_raw", C.c_void_p), ("scaling_act", C.c_void_p), ("rotation_act", C.c_void_p)]


class ExpandGrads(C.Structure):
    _fields_ = [("dL_dxyz", C.c_void_p), ("dL_dscaling_log", C.c_void_p), ("dL_drotation_raw", C.c_void_p),
                ("dL_dscaling_act", C.c_void_p), ("dL_drotation_act", C.c_void_p),
                ("dL_dvertices", C.c_void_p), ("dL_dtriangles", C.c_void_p), ("dL_dalpha_raw", C.c_void_p),
                ("dL_dscale_raw", C.c_void_p)]


class PointsArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("triangles", C.c_void_p), ("eps", C.c_float), ("xyz", C.c_void_p),
                ("scaling_log", C.c_void_p), ("rotation_raw", C.c_void_p), ("scaling_act", C.c_void_p),
                ("rotation_act", C.c_void_p)]


class PointsVerticesArgs(C.Structure):
    """struct gms_points_vertices_args"""
    _fields_ = [("P", C.c_int32), ("xyz", C.c_void_p), ("scaling_log", C.c_void_p), ("scaling_cols", C.c_int32),
                ("rotation_raw", C.c_void_p), ("triangles", C.c_void_p)]


class LossArgs(C.Structure):
    _fields_ = [("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("img", C.c_void_p), ("gt", C.c_void_p),
                ("lambda_dssim", C.c_float), ("dL_dloss", C.c_void_p), ("loss", C.c_void_p), ("dL_dimg", C.c_void_p),
                ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t)]


class AdamArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("offset", C.c_int64), ("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("nseg", C.c_int32), ("seg_end", C.c_int64 * 8), ("lr0", C.c_double * 8), ("lr1", C.c_double * 8),
                ("inner", C.c_int32 * 8), ("period", C.c_int32 * 8), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("step", C.c_int32), ("zero_grad", C.c_int32), ("zero_end", C.c_int64)]


class FrameArgs(C.Structure):
    _fields_ = [("V", C.c_int32), ("F", C.c_int32), ("K", C.c_int32), ("M", C.c_int32),
                ("vertices", C.c_void_p), ("faces", C.c_void_p), ("alpha_raw", C.c_void_p), ("scale_raw", C.c_void_p),
                ("features", C.c_void_p), ("opacity_raw", C.c_void_p), ("eps", C.c_float),
                ("d_vertices", C.c_void_p), ("d_alpha_raw", C.c_void_p), ("d_scale_raw", C.c_void_p),
                ("d_features", C.c_void_p), ("d_opacity_raw", C.c_void_p),
                ("settings", RasterSettings), ("gt", C.c_void_p), ("lambda_dssim", C.c_float), ("loss", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("num_rendered", C.POINTER(C.c_int64)),
                ("binning_capacity", C.c_int64), ("n_host_mapped", C.c_void_p), ("d_color_sh", C.c_void_p),
                ("event_sh_ready", C.c_void_p), ("event_loss_ready", C.c_void_p)]


class FrameView(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p), ("opacities", C.c_void_p),
                ("radii", C.c_void_p), ("image", C.c_void_p), ("invdepth", C.c_void_p)]


class AdamShArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("M", C.c_int32), ("sh_degree", C.c_int32), ("R", C.c_int32), ("xyz", C.c_void_p),
                ("exchange", C.c_void_p), ("slot_floats", C.c_int64), ("grad_scale", C.c_float), ("p", C.c_void_p),
                ("m", C.c_void_p), ("v", C.c_void_p), ("lr_dc", C.c_double), ("lr_rest", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("step", C.c_int32)]


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)

# every symbol include/gms_b200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = ["gms_scratch_bytes", "gms_binning_bytes", "gms_rasterize_forward", "gms_rasterize_forward_nosync", "gms_rasterize_backward",
               "gms_mark_visible", "gms_debug_get_views", "gms_debug_unpack", "gms_expand_forward",
               "gms_expand_backward", "gms_last_error", "gms_version", "gms_launch_count", "gms_set_option",
               "gms_kernel_times", "gms_loss_scratch_bytes", "gms_l1_ssim_loss", "gms_adam_step",
               "gms_frame_workspace_bytes", "gms_train_frame", "gms_points_expand_forward",
               "gms_points_prepare_vertices", "gms_image_quantize", "gms_image_dequantize", "gms_adam_sh_factored", "gms_frame_views"]

_lib = None


def lib():
    """Load libgms_b200.so; fail loudly if it is missing (no fallback path exists by design)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GmsLibraryError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run "
            f"`python gaussian-mesh-splatting_b200/build.py` (needs nvcc, sm_100a). There is no CPU fallback.")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise GmsLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    L.gms_last_error.restype = C.c_char_p
    L.gms_version.restype = C.c_char_p
    L.gms_launch_count.restype = C.c_int64
    L.gms_launch_count.argtypes = [C.c_int]
    L.gms_binning_bytes.restype = C.c_size_t
    L.gms_binning_bytes.argtypes = [C.c_int64, C.c_int32]
    L.gms_scratch_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.gms_rasterize_forward.argtypes = [C.POINTER(RasterSettings), C.POINTER(RasterInputs), C.POINTER(RasterOutputs),
                                        ALLOC_FN, C.c_void_p, C.POINTER(RasterSaved), C.c_void_p]
    L.gms_rasterize_forward_nosync.argtypes = [C.POINTER(RasterSettings), C.POINTER(RasterInputs), C.POINTER(RasterOutputs),
                                               ALLOC_FN, C.c_void_p, C.POINTER(RasterSaved), C.c_int64, C.c_void_p, C.c_void_p]
    L.gms_rasterize_backward.argtypes = [C.POINTER(RasterSettings), C.POINTER(RasterInputs), C.c_void_p,
                                         C.POINTER(RasterSaved), C.c_void_p, C.c_void_p, C.POINTER(RasterGrads),
                                         C.c_void_p]
    L.gms_mark_visible.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gms_debug_get_views.argtypes = [C.POINTER(RasterSaved), C.c_int32, C.c_int32, C.c_int32, C.POINTER(DebugViews)]
    L.gms_debug_unpack.argtypes = [C.POINTER(RasterSaved), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.gms_expand_forward.argtypes = [C.POINTER(ExpandArgs), C.c_void_p]
    L.gms_expand_backward.argtypes = [C.POINTER(ExpandArgs), C.POINTER(ExpandGrads), C.c_void_p]
    L.gms_set_option.argtypes = [C.c_char_p, C.c_int]
    L.gms_loss_scratch_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    L.gms_l1_ssim_loss.argtypes = [C.POINTER(LossArgs), C.c_void_p]
    L.gms_adam_step.argtypes = [C.POINTER(AdamArgs), C.c_void_p]
    L.gms_points_expand_forward.argtypes = [C.POINTER(PointsArgs), C.c_void_p]
    L.gms_points_prepare_vertices.argtypes = [C.POINTER(PointsVerticesArgs), C.c_void_p]
    L.gms_image_quantize.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.gms_image_dequantize.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.gms_adam_sh_factored.argtypes = [C.POINTER(AdamShArgs), C.c_void_p]
    L.gms_frame_views.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(FrameView)]
    L.gms_frame_workspace_bytes.restype = C.c_size_t
    L.gms_frame_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.gms_train_frame.argtypes = [C.POINTER(FrameArgs), ALLOC_FN, C.c_void_p, C.c_void_p]
    _lib = L
    # GMS_OPTIONS="key=value,key=value": tuning knobs applied at load (A/B runs of whole test suites / benches)
    for kv in filter(None, os.environ.get("GMS_OPTIONS", "").split(",")):
        k, _, v = kv.partition("=")
        if L.gms_set_option(k.strip().encode(), int(v)) == -1:
            raise GmsLibraryError(f"GMS_OPTIONS: unknown option {k!r}")
    return L


def check(rc: int, what: str) -> None:
    if rc != GMS_OK:
        msg = lib().gms_last_error().decode("utf-8", "replace")
        if rc == GMS_E_ARG:
            # same exception type/text as the stock Python shim for bad argument combinations
            raise Exception(msg)
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def set_option(key: str, value: int) -> int:
    return int(lib().gms_set_option(key.encode(), int(value)))


def kernel_times(reset: bool = True) -> dict:
    """{kernel name: (accumulated ms, launches)} measured by CUDA events on the launching stream."""
    L = lib()
    n = 16
    ms = (C.c_double * n)(); cnt = (C.c_int64 * n)(); names = (C.c_char_p * n)()
    k = L.gms_kernel_times(1 if reset else 0, n, ms, cnt, names)
    return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(min(k, n)) if names[i]}


def launch_count(reset: bool = False) -> int:
    return int(lib().gms_launch_count(1 if reset else 0))
