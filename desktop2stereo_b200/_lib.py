"""ctypes binding of libd2s_b200.so (the C ABI declared in include/d2s_b200.h).

The library is the product: if it is missing or a call fails this module raises — there is no
CPU or PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libd2s_b200.so")

F32, F16, BF16, U8 = 0, 1, 2, 3
FULL_SBS, HALF_SBS, FULL_TAB, HALF_TAB = 0, 1, 2, 3
DISPLAY_MODES = {"Full-SBS": FULL_SBS, "Half-SBS": HALF_SBS, "Full-TAB": FULL_TAB, "Half-TAB": HALF_TAB}
WARP_BILINEAR, WARP_GATHER = 0, 1


class D2SError(RuntimeError):
    pass


class Image(C.Structure):
    _fields_ = [("base", C.c_void_p), ("dtype", C.c_int32), ("reserved", C.c_int32),
                ("sc", C.c_int64), ("sy", C.c_int64), ("sx", C.c_int64)]


class WarpParams(C.Structure):
    _fields_ = [("rgb", Image), ("out", Image), ("depth", C.c_void_p), ("depth_dtype", C.c_int32),
                ("depth_h", C.c_int32), ("depth_w", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("ipd_uv", C.c_double), ("depth_ratio", C.c_double), ("convergence", C.c_double),
                ("display_mode", C.c_int32), ("fill_16_9", C.c_int32), ("warp_mode", C.c_int32),
                ("rgb_round_to_depth_dtype", C.c_int32),
                ("idx_left", C.c_void_p), ("idx_right", C.c_void_p)]


class ModelConfig(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32), ("mlp_hidden", C.c_int32),
                ("patch", C.c_int32), ("pos_grid", C.c_int32), ("out_indices", C.c_int32 * 4),
                ("neck", C.c_int32 * 4), ("fusion", C.c_int32), ("head_hidden", C.c_int32),
                ("layer_norm_eps", C.c_float), ("max_depth", C.c_float), ("metric", C.c_int32),
                ("max_batch", C.c_int32), ("max_h", C.c_int32), ("max_w", C.c_int32),
                ("temporal", C.c_int32), ("pos_interp_offset", C.c_float), ("reserved", C.c_int32 * 2)]


class PostParams(C.Structure):
    _fields_ = [("depth_in", C.c_void_p), ("in_dtype", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("out", C.c_void_p), ("out_dtype", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("compute_dtype", C.c_int32), ("metric", C.c_int32), ("percentile", C.c_float),
                ("subsample_cap", C.c_int32), ("gamma", C.c_float), ("foreground_scale", C.c_float),
                ("aa_strength", C.c_float), ("ema_state", C.c_void_p), ("ema_valid", C.c_int32),
                ("ema_alpha", C.c_float), ("out_lowres", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class DibrParams(C.Structure):
    _fields_ = [("rgb", Image), ("out", Image), ("depth", C.c_void_p), ("depth_dtype", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("display_mode", C.c_int32), ("ipd_uv", C.c_double), ("depth_ratio", C.c_double), ("convergence", C.c_double),
                ("roll", C.c_double), ("resolution_x", C.c_float), ("resolution_y", C.c_float), ("search_radius", C.c_int32),
                ("depth_tolerance", C.c_float), ("blur_radius", C.c_float), ("feather_enabled", C.c_int32),
                ("feather_width", C.c_float), ("corner_radius", C.c_float), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class PipeConfig(C.Structure):
    _fields_ = [("frame_h", C.c_int32), ("frame_w", C.c_int32), ("channels", C.c_int32), ("target_height", C.c_int32),
                ("rgb_dtype", C.c_int32), ("depth_resolution", C.c_int32), ("patch", C.c_int32),
                ("mean", C.c_float * 3), ("std", C.c_float * 3), ("metric", C.c_int32), ("percentile", C.c_float),
                ("subsample_cap", C.c_int32), ("gamma", C.c_float), ("foreground_scale", C.c_float), ("aa_strength", C.c_float),
                ("use_temporal_smooth", C.c_int32), ("ema_alpha", C.c_float),
                ("ipd_uv", C.c_double), ("depth_ratio", C.c_double), ("convergence", C.c_double),
                ("display_mode", C.c_int32), ("fill_16_9", C.c_int32), ("out_dtype", C.c_int32), ("out_format", C.c_int32),
                ("slots", C.c_int32), ("host_io", C.c_int32), ("streams", C.c_int32),
                ("jpeg_quality", C.c_int32), ("jpeg_restart_interval", C.c_int32), ("fps_overlay", C.c_int32)]


# every symbol include/d2s_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "d2s_sbs_out_shape": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "d2s_make_sbs": (C.c_int, [C.POINTER(WarpParams), C.c_void_p]),
    "d2s_process": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "d2s_model_input_shape": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "d2s_preprocess": (C.c_int, [C.POINTER(Image), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_size_t, C.c_void_p]),
    "d2s_preprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "d2s_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(ModelConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "d2s_destroy": (C.c_int, [C.c_void_p]),
    "d2s_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "d2s_set_policy": (C.c_int, [C.c_void_p, C.c_int]),
    "d2s_reset_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "d2s_release_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "d2s_debug_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]),
    "d2s_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "d2s_launch_count": (C.c_int64, []),
    "d2s_postprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "d2s_postprocess": (C.c_int, [C.POINTER(PostParams), C.c_void_p]),
    "d2s_overlay_fps": (C.c_int, [C.POINTER(Image), C.c_int, C.c_int, C.c_char_p, C.c_void_p]),
    "d2s_rgb_to_nv12": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "d2s_jpeg_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "d2s_jpeg_max_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "d2s_jpeg_encode": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "d2s_dibr_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "d2s_dibr_out_shape": (C.c_int, [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "d2s_make_sbs_dibr": (C.c_int, [C.POINTER(DibrParams), C.c_void_p]),
    "d2s_pipe_create": (C.c_int, [C.c_void_p, C.POINTER(PipeConfig), C.POINTER(C.c_void_p)]),
    "d2s_pipe_destroy": (C.c_int, [C.c_void_p]),
    "d2s_pipe_set_fps_text": (C.c_int, [C.c_void_p, C.c_char_p]),
    "d2s_pipe_geometry": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 6 + [C.POINTER(C.c_size_t)] * 2),
    "d2s_pipe_slot_buffers": (C.c_int, [C.c_void_p, C.c_int] + [C.POINTER(C.c_void_p)] * 6),
    "d2s_pipe_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "d2s_pipe_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "d2s_pipe_reset": (C.c_int, [C.c_void_p]),
    "d2s_pipe_set_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "d2s_pipe_slot_times": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "d2s_debug_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "d2s_debug_set_gemm_policy": (C.c_int, [C.c_int]),
    "d2s_debug_conv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "d2s_debug_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "d2s_debug_force_generic_warp": (C.c_int, [C.c_int]),
    "d2s_last_error": (C.c_char_p, []),
    "d2s_version": (C.c_char_p, []),
}

_lib = None


def lib():
    """Load libd2s_b200.so (built by `python -m desktop2stereo_b200.build`).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise D2SError(f"{LIB_PATH} not built: run `python -m desktop2stereo_b200.build` "
                           "(there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().d2s_last_error().decode("utf-8", "replace")
        raise D2SError(f"{what or 'd2s call'} failed (status {rc}): {msg}")
