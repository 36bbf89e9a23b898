"""ORACLE (test infrastructure) — ctypes front-end of `warp_oracle.c`, plus a torch restatement.

`make_sbs_core_oracle`  : strict-fp32 C restatement of depth.py:2122-2184 (both warp branches).
`make_sbs_core_torch`   : the same function restated with the reference's own torch ops
                          (linspace + grid_sample / gather), device-agnostic.  On the GPU box it runs
                          on CUDA and so exercises the very ATen kernels the reference would call.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import build

DT = {"float32": 0, "float16": 1, "bfloat16": 2}
MODES = {"Full-SBS": 0, "Half-SBS": 1, "Full-TAB": 2, "Half-TAB": 3}
_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        f = ctypes.c_float
        _lib.d2s_oracle_make_sbs.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            f, f, f, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.d2s_oracle_make_sbs.restype = ctypes.c_int
        _lib.d2s_oracle_out_shape.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2
        _lib.d2s_oracle_linspace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return _lib


def out_shape(h, w, display_mode="Half-SBS", fill_16_9=False):
    oh, ow = ctypes.c_int(), ctypes.c_int()
    _load().d2s_oracle_out_shape(h, w, MODES[display_mode], int(fill_16_9), ctypes.byref(oh), ctypes.byref(ow))
    return oh.value, ow.value


def linspace(n, fma=True):
    out = np.empty(n, np.float32)
    _load().d2s_oracle_linspace(out.ctypes.data, n, int(fma))
    return out


def make_sbs_core_oracle(rgb, depth, ipd_uv=0.064, depth_ratio=2.0, display_mode="Half-SBS",
                         fill_16_9=False, convergence=0.0, depth_dtype="float32", gather=False,
                         fma=True, xs=None, ys=None, return_indices=False, out_dtype=None):
    """rgb [3,h,w] float array (values of the eye dtype), depth [h,w] float array holding values of
    `depth_dtype`.  Returns out [3,oh,ow] float32 (and int32 index maps if asked)."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    _, h, w = rgb.shape
    oh, ow = out_shape(h, w, display_mode, fill_16_9)
    out = np.empty((3, oh, ow), np.float32)
    il = np.empty((h, w), np.int32)
    ir = np.empty((h, w), np.int32)
    xs_p = None if xs is None else np.ascontiguousarray(xs, np.float32)
    ys_p = None if ys is None else np.ascontiguousarray(ys, np.float32)
    rc = _load().d2s_oracle_make_sbs(
        rgb.ctypes.data, depth.ctypes.data, DT[depth_dtype], h, w,
        float(np.float32(ipd_uv * w)), float(depth_ratio), float(convergence),
        MODES[display_mode], int(fill_16_9), int(gather), int(fma),
        DT[out_dtype] if out_dtype else (DT[depth_dtype] if gather else 0),
        None if xs_p is None else xs_p.ctypes.data, None if ys_p is None else ys_p.ctypes.data,
        out.ctypes.data, il.ctypes.data, ir.ctypes.data)
    assert rc == 0
    return (out, il, ir) if return_indices else out


def make_sbs_core_torch(rgb, depth, ipd_uv=0.064, depth_ratio=2.0, display_mode="Half-SBS",
                        fill_16_9=False, convergence=0.0, gather=False):
    """Restatement of depth.py:2122-2184 with the reference's own torch calls, no autocast context
    (grid_sample is in autocast's fp32 list, so inputs are promoted to fp32 explicitly here)."""
    import torch
    import torch.nn.functional as F
    device = rgb.device
    C, H, W = rgb.shape
    img = rgb.unsqueeze(0).clamp(0, 255)
    depth = depth - convergence
    inv = -depth * depth_ratio
    max_px = ipd_uv * W
    shifts = inv * max_px * 0.05
    if not gather:
        xs = torch.linspace(-1.0, 1.0, W, device=device).view(1, 1, W).expand(1, H, W)
        ys = torch.linspace(-1.0, 1.0, H, device=device).view(1, H, 1).expand(1, H, W)
        shift_norm = shifts * (2.0 / (W - 1))
        gl = torch.stack([xs + shift_norm, ys], dim=-1)
        gr = torch.stack([xs - shift_norm, ys], dim=-1)
        left = F.grid_sample(img.float(), gl.float(), mode="bilinear", padding_mode="reflection", align_corners=True)[0]
        right = F.grid_sample(img.float(), gr.float(), mode="bilinear", padding_mode="reflection", align_corners=True)[0]
    else:
        base = torch.arange(W, device=device, dtype=torch.int64).view(1, -1).expand(H, -1)
        shifts = shifts.to(torch.float32)
        cl = (base.to(torch.float32) + shifts).clamp(0, W - 1).long()
        cr = (base.to(torch.float32) - shifts).clamp(0, W - 1).long()
        left = torch.gather(img.expand(1, C, H, W), 3, cl.unsqueeze(0).expand(C, H, W).unsqueeze(0))[0]
        right = torch.gather(img.expand(1, C, H, W), 3, cr.unsqueeze(0).expand(C, H, W).unsqueeze(0))[0]
    if fill_16_9:
        left, right = _pad_16_9(left), _pad_16_9(right)
    out = torch.cat([left, right], dim=1 if display_mode in ("Half-TAB", "Full-TAB") else 2)
    if display_mode not in ("Full-SBS", "Full-TAB"):
        out = F.interpolate(out.unsqueeze(0), size=left.shape[1:], mode="area")[0]
    return out.clamp(0, 255)


def _pad_16_9(t):
    import torch.nn.functional as F
    _, h, w = t.shape
    r_img, r_t = w / h, 16 / 9
    if abs(r_img - r_t) < 1e-3:
        return t
    if r_img > r_t:
        nh = int(round(w / r_t)); top = (nh - h) // 2
        return F.pad(t, (0, 0, top, nh - h - top))
    nw = int(round(h * r_t)); left = (nw - w) // 2
    return F.pad(t, (left, nw - w - left, 0, 0))
