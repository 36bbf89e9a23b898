"""ORACLE (test infrastructure) — ctypes front-end of `dibr_oracle.c` (the reference viewer's DIBR fragment shader,
viewer.py:386-631, evaluated per output pixel on the CPU) and the packing of the two eye views the viewer's viewports do
(viewer.py:2680-2760).  PARITY UNPINNED against the reference (no OpenGL context here; the reference never sets u_resolution) —
see the C file's header; cross-checked against an independent numpy restatement in tests/test_oracle_dibr.py."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        f, i, p = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
        _lib.d2s_oracle_dibr.argtypes = [p, p, i, i, i, i, f, f, f, f, f, f, f, i, f, f, p, p, i, f, f, p, p, p, p]
        _lib.d2s_oracle_dibr.restype = i
    return _lib


def view_shape(h, w, display_mode):
    vh = h // 2 if display_mode == "Half-TAB" else h
    vw = w // 2 if display_mode == "Half-SBS" else w
    return vh, vw


def exp_tables(n=32):
    i = np.arange(n + 1, dtype=np.float64)
    return np.exp(-i * 0.15).astype(np.float32), np.exp(-i * 0.2).astype(np.float32)


def eye_views(rgb_u8_hwc, depth, ipd_uv=0.064, depth_ratio=1.0, convergence=0.0, display_mode="Half-SBS", roll=0.0, resolution=None,
              search_radius=12, depth_tolerance=0.012, blur_radius=2.5, feather_enabled=False, feather_width=0.0, corner_radius=0.0,
              return_conf=False):
    """rgb [h,w,3] (values 0..255), depth [h,w] float32 -> (left, right) rgba float32 [vh,vw,4] (colour in [0,1], alpha)"""
    h, w = depth.shape
    vh, vw = view_shape(h, w, display_mode)
    color = np.ascontiguousarray(rgb_u8_hwc.astype(np.float32) / np.float32(255.0))       # the normalised u8 texture
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    res = resolution if resolution is not None else (float(vw), float(vh))
    w1, w2 = exp_tables()
    left, right = np.empty((vh, vw, 4), np.float32), np.empty((vh, vw, 4), np.float32)
    cl, cr = np.empty((vh, vw), np.float32), np.empty((vh, vw), np.float32)
    rc = _load().d2s_oracle_dibr(color.ctypes.data, depth.ctypes.data, h, w, vw, vh, res[0], res[1], ipd_uv,
                                 np.float32(0.1 * depth_ratio), convergence, np.float32(math.cos(roll)), np.float32(math.sin(roll)),
                                 search_radius, depth_tolerance, blur_radius, w1.ctypes.data, w2.ctypes.data,
                                 int(feather_enabled), feather_width, corner_radius,
                                 left.ctypes.data, right.ctypes.data, cl.ctypes.data, cr.ctypes.data)
    assert rc == 0
    return (left, right, cl, cr) if return_conf else (left, right)


def make_sbs_dibr_oracle(rgb_u8_hwc, depth, display_mode="Half-SBS", **kw):
    """The packed frame [oh, ow, 3] float32, 0..255: colour x alpha over black, eyes laid out like the viewer's viewports."""
    left, right = eye_views(rgb_u8_hwc, depth, display_mode=display_mode, **kw)

    def present(v):
        c = v[..., :3] * v[..., 3:4]
        return np.minimum(np.maximum(c, np.float32(0)), np.float32(1)) * np.float32(255)
    axis = 0 if display_mode in ("Full-TAB", "Half-TAB") else 1
    return np.concatenate([present(left), present(right)], axis=axis)


def synthetic_scene(seed, h, w):
    """a frame with a near object in front of a far background: produces real disocclusions on both sides of the object"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    rgb = np.stack([(xx * 3 + yy * 2 + 40 * c) % 256 for c in range(3)], -1).astype(np.uint8)
    box = rgb[h // 4: 3 * h // 4, w // 3: 2 * w // 3]
    box[...] = rng.integers(0, 256, box.shape)
    depth = np.full((h, w), 0.15, np.float32) + 0.1 * (yy / h).astype(np.float32)
    depth[h // 4: 3 * h // 4, w // 3: 2 * w // 3] = 0.9
    return rgb, depth.astype(np.float32)
