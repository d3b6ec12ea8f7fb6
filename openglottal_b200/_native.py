"""ctypes binding of libopenglottal_b200.so (the C ABI in include/openglottal_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every compute
call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import LIB_PATH

DTYPE_U8, DTYPE_F32 = 0, 1
PRECISION_BF16, PRECISION_F32, PRECISION_F16 = 0, 1, 2

_c_float_p = C.POINTER(C.c_float)


class ConvBN(C.Structure):
    _fields_ = [
        ("weight", _c_float_p),
        ("bn_weight", _c_float_p),
        ("bn_bias", _c_float_p),
        ("running_mean", _c_float_p),
        ("running_var", _c_float_p),
    ]


class ConvT(C.Structure):
    _fields_ = [("weight", _c_float_p), ("bias", _c_float_p)]


class UNetState(C.Structure):
    _fields_ = [
        ("downs", (ConvBN * 2) * 4),
        ("bottleneck", ConvBN * 2),
        ("up_t", ConvT * 4),
        ("up_c", (ConvBN * 2) * 4),
        ("head_weight", _c_float_p),
        ("head_bias", _c_float_p),
        ("bn_eps", C.c_float),
    ]


# every symbol include/openglottal_b200.h declares
EXPORTS = [
    "ogl_version",
    "ogl_last_error",
    "ogl_unet_create",
    "ogl_unet_destroy",
    "ogl_unet_load_state",
    "ogl_unet_prepare",
    "ogl_unet_workspace_bytes",
    "ogl_unet_forward",
    "ogl_unet_set_profiling",
    "ogl_unet_layer_times",
    "ogl_unet_launch_count",
    "ogl_unet_launch_name",
    "ogl_unet_set_schedule",
    "ogl_unet_set_cta_pairs",
    "ogl_unet_set_repeat",
    "ogl_unet_set_fused_stem",
    "ogl_unet_set_compose",
    "ogl_features_workspace_bytes",
    "ogl_features",
    "ogl_features_f64",
    "ogl_bgr_to_gray",
    "ogl_mask_area_boxes",
    "ogl_letterbox_crops",
    "ogl_unletterbox_area",
    "ogl_mask_overlap_counts",
    "ogl_resize_u8_linear",
    "ogl_prob_resize_mask",
    "ogl_debug_tc_layer",
    "ogl_debug_s2d_layer",
    "ogl_debug_upcat_layer",
    "ogl_debug_upcat_program",
    "ogl_debug_s2d_program",
]

_lib = None


def library_path() -> Path:
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (built by ``__graft_entry__.build()`` / ``build.build_library``)."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    lib_path = Path(os.environ.get("OGL_LIB", str(LIB_PATH)))   # experiment builds
    if not lib_path.exists():
        raise RuntimeError(
            f"{lib_path} is missing: run `python -m openglottal_b200.build` (needs nvcc). "
            "openglottal_b200 has no CPU or PyTorch fallback."
        )
    _lib = load_path(lib_path)
    return _lib


def load_path(lib_path) -> C.CDLL:
    """Bind the C ABI of the library at ``lib_path`` (not cached: same-process comparisons of two
    builds, scripts/byte_ops_bench.py)."""
    lib = C.CDLL(str(lib_path))
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    lib.ogl_version.restype = i32
    lib.ogl_version.argtypes = []
    lib.ogl_last_error.restype = C.c_char_p
    lib.ogl_last_error.argtypes = []
    lib.ogl_unet_create.restype = i32
    lib.ogl_unet_create.argtypes = [C.POINTER(vp), i32]
    lib.ogl_unet_destroy.restype = i32
    lib.ogl_unet_destroy.argtypes = [vp]
    lib.ogl_unet_load_state.restype = i32
    lib.ogl_unet_load_state.argtypes = [vp, C.POINTER(UNetState)]
    lib.ogl_unet_prepare.restype = i32
    lib.ogl_unet_prepare.argtypes = [vp, i32]
    lib.ogl_unet_workspace_bytes.restype = sz
    lib.ogl_unet_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
    lib.ogl_unet_forward.restype = i32
    lib.ogl_unet_forward.argtypes = [vp, vp, i32, i32, i32, i32, vp, sz, vp, vp, vp, C.c_float,
                                     i32, vp]
    lib.ogl_unet_set_profiling.restype = i32
    lib.ogl_unet_set_profiling.argtypes = [vp, i32]
    lib.ogl_unet_layer_times.restype = i32
    lib.ogl_unet_layer_times.argtypes = [vp, _c_float_p, i32, C.POINTER(i32)]
    lib.ogl_unet_launch_count.restype = i32
    lib.ogl_unet_launch_count.argtypes = [vp]
    lib.ogl_unet_launch_name.restype = C.c_char_p
    lib.ogl_unet_launch_name.argtypes = [vp, i32]
    lib.ogl_unet_set_fused_stem.restype = i32
    lib.ogl_unet_set_fused_stem.argtypes = [vp, i32]
    lib.ogl_unet_set_compose.restype = i32
    lib.ogl_unet_set_compose.argtypes = [vp, i32]
    lib.ogl_unet_set_cta_pairs.restype = i32
    lib.ogl_unet_set_cta_pairs.argtypes = [vp, i32]
    lib.ogl_unet_set_repeat.restype = i32
    lib.ogl_unet_set_repeat.argtypes = [vp, i32, i32]
    lib.ogl_unet_set_schedule.restype = i32
    lib.ogl_unet_set_schedule.argtypes = [vp, i32]
    lib.ogl_features_workspace_bytes.restype = sz
    lib.ogl_features_workspace_bytes.argtypes = [i64]
    lib.ogl_features.restype = i32
    lib.ogl_features.argtypes = [vp, i64, vp, vp, vp, sz, vp]
    lib.ogl_features_f64.restype = i32
    lib.ogl_features_f64.argtypes = [vp, i64, vp, vp, vp, sz, vp]
    lib.ogl_bgr_to_gray.restype = i32
    lib.ogl_bgr_to_gray.argtypes = [vp, vp, i64, vp]
    lib.ogl_mask_area_boxes.restype = i32
    lib.ogl_mask_area_boxes.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.ogl_letterbox_crops.restype = i32
    lib.ogl_letterbox_crops.argtypes = [vp, i32, i32, i32, vp, i32, vp, vp]
    lib.ogl_unletterbox_area.restype = i32
    lib.ogl_unletterbox_area.argtypes = [vp, i32, i32, vp, i32, i32, vp, vp, vp]
    lib.ogl_mask_overlap_counts.restype = i32
    lib.ogl_mask_overlap_counts.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.ogl_resize_u8_linear.restype = i32
    lib.ogl_resize_u8_linear.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp]
    lib.ogl_prob_resize_mask.restype = i32
    lib.ogl_prob_resize_mask.argtypes = [vp, i32, i32, i32, i32, i32, C.c_float, vp, vp, vp]
    lib.ogl_debug_tc_layer.restype = i32
    lib.ogl_debug_tc_layer.argtypes = [vp, i32, vp, i32, vp, i32, _c_float_p, _c_float_p, i32,
                                       i32, i32, i32, vp, vp, vp]
    lib.ogl_debug_s2d_layer.restype = i32
    lib.ogl_debug_s2d_layer.argtypes = [vp, i32, vp, i32, vp, _c_float_p, _c_float_p, _c_float_p,
                                        _c_float_p, i32, i32, i32, vp, vp, vp]
    lib.ogl_debug_upcat_layer.restype = i32
    lib.ogl_debug_upcat_layer.argtypes = [vp, vp, vp, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                          i32, i32, i32, i32, vp, vp]
    lib.ogl_debug_upcat_program.restype = i32
    lib.ogl_debug_upcat_program.argtypes = [_c_float_p, _c_float_p, _c_float_p, _c_float_p, i32, vp, vp,
                                            vp, vp, vp]
    lib.ogl_debug_s2d_program.restype = i32
    lib.ogl_debug_s2d_program.argtypes = [_c_float_p, _c_float_p, i32, _c_float_p, _c_float_p, vp,
                                          sz, C.POINTER(sz), vp, i32, C.POINTER(i32), vp,
                                          C.POINTER(i32), vp]
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ogl_last_error()
        raise RuntimeError("openglottal_b200: " + (msg.decode() if msg else f"error {rc}"))
