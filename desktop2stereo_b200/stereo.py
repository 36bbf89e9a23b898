"""Host-side mirror of the reference's stereo functions, backed by libd2s_b200's warp kernel.

Same names, argument meaning and defaults as the reference:
    make_sbs_core   depth.py:2122-2184
    make_sbs        depth.py:2186-2231
PyTorch appears only as the owner of device memory and of the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DISPLAY_MODES, Image, WarpParams

_TORCH2D2S = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.uint8: _lib.U8}
_D2S2TORCH = {v: k for k, v in _TORCH2D2S.items()}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.D2SError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU path")


def image_view(t: torch.Tensor, layout: str) -> Image:
    """Describe a torch tensor as a d2s_image.  layout: 'CHW' [3,h,w], 'HWC' [h,w,3] (RGB) or
    'BGRA'/'BGR' [h,w,4|3] (capture order, read as RGB through a negative channel stride)."""
    es = t.element_size()
    img = Image()
    img.dtype = _TORCH2D2S[t.dtype]
    if layout == "CHW":
        assert t.dim() == 3 and t.shape[0] >= 3
        img.base, img.sc, img.sy, img.sx = t.data_ptr(), t.stride(0), t.stride(1), t.stride(2)
    elif layout == "HWC":
        assert t.dim() == 3 and t.shape[2] >= 3
        img.base, img.sc, img.sy, img.sx = t.data_ptr(), t.stride(2), t.stride(0), t.stride(1)
    elif layout in ("BGRA", "BGR"):
        assert t.dim() == 3 and t.shape[2] >= 3
        img.base = t.data_ptr() + 2 * t.stride(2) * es
        img.sc, img.sy, img.sx = -t.stride(2), t.stride(0), t.stride(1)
    else:
        raise ValueError(layout)
    return img


def sbs_out_shape(h: int, w: int, display_mode: str = "Half-SBS", fill_16_9: bool = False):
    oh, ow = C.c_int(), C.c_int()
    _lib.check(_lib.lib().d2s_sbs_out_shape(h, w, DISPLAY_MODES[display_mode], int(fill_16_9),
                                            C.byref(oh), C.byref(ow)), "d2s_sbs_out_shape")
    return oh.value, ow.value


def make_sbs_core(rgb: torch.Tensor, depth: torch.Tensor, ipd_uv=0.064, depth_ratio=2.0,
                  display_mode="Half-SBS", fill_16_9=False, convergence=0.0, device=None, *,
                  rgb_layout="CHW", out: torch.Tensor | None = None, out_dtype=None, out_layout="CHW",
                  gather=False, return_indices=False, round_rgb_to_depth_dtype=False):
    """depth.py:2122-2184.  rgb [3,h,w] (0..255), depth [h,w] in [0,1] (or a lower-resolution map,
    which is bilinearly upsampled inside the kernel exactly as depth.py:1998-2004 would).  Returns the
    packed stereo image [3,oh,ow] (or [oh,ow,3] with out_layout='HWC'), fp32 like the reference's
    grid_sample branch unless out_dtype says otherwise."""
    _require_cuda(rgb, "rgb"), _require_cuda(depth, "depth")
    if display_mode not in DISPLAY_MODES:
        raise ValueError(f"display_mode {display_mode!r}")
    if rgb_layout == "CHW":
        h, w = rgb.shape[1:]
    else:
        h, w = rgb.shape[:2]
    if depth.dim() != 2:
        depth = depth.squeeze()
    depth = depth.contiguous()
    oh, ow = sbs_out_shape(h, w, display_mode, fill_16_9)
    if out is None:
        if out_dtype is None:
            out_dtype = depth.dtype if gather else torch.float32
        shape = (3, oh, ow) if out_layout == "CHW" else (oh, ow, 3)
        out = torch.empty(shape, dtype=out_dtype, device=rgb.device)
    p = WarpParams()
    p.rgb = image_view(rgb, rgb_layout)
    p.out = image_view(out, out_layout)
    p.depth, p.depth_dtype = depth.data_ptr(), _TORCH2D2S[depth.dtype]
    p.depth_h, p.depth_w, p.h, p.w = depth.shape[0], depth.shape[1], h, w
    p.ipd_uv, p.depth_ratio, p.convergence = float(ipd_uv), float(depth_ratio), float(convergence)
    p.display_mode, p.fill_16_9 = DISPLAY_MODES[display_mode], int(bool(fill_16_9))
    p.warp_mode = _lib.WARP_GATHER if gather else _lib.WARP_BILINEAR
    p.rgb_round_to_depth_dtype = int(bool(round_rgb_to_depth_dtype))
    il = ir = None
    if return_indices:
        il = torch.empty((h, w), dtype=torch.int32, device=rgb.device)
        ir = torch.empty((h, w), dtype=torch.int32, device=rgb.device)
        p.idx_left, p.idx_right = il.data_ptr(), ir.data_ptr()
    with torch.cuda.device(rgb.device):
        _lib.check(_lib.lib().d2s_make_sbs(C.byref(p), _stream_ptr(rgb.device)), "d2s_make_sbs")
    return (out, il, ir) if return_indices else out


def rgb_to_nv12(rgb_hwc: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Packed u8 HWC frame [h,w,3] -> NV12 [h*3//2, w] u8 (libjpeg's RGB->YCbCr + 4:2:0 stages, what cv2.imencode does first on
    the frame the reference streams, streamer.py:250-256).  h, w even."""
    _require_cuda(rgb_hwc, "rgb")
    if rgb_hwc.dtype != torch.uint8 or rgb_hwc.dim() != 3 or rgb_hwc.shape[2] != 3 or rgb_hwc.stride(2) != 1 or rgb_hwc.stride(1) != 3:
        raise ValueError("rgb_to_nv12 expects a uint8 [h,w,3] tensor with packed pixels")
    h, w, _ = rgb_hwc.shape
    if out is None:
        out = torch.empty((h * 3 // 2, w), dtype=torch.uint8, device=rgb_hwc.device)
    with torch.cuda.device(rgb_hwc.device):
        _lib.check(_lib.lib().d2s_rgb_to_nv12(rgb_hwc.data_ptr(), rgb_hwc.stride(0), h, w, out.data_ptr(), _stream_ptr(rgb_hwc.device)), "d2s_rgb_to_nv12")
    return out


class JpegEncoder:
    """Device-side baseline JPEG encoder for packed u8 HWC frames of one size (d2s_jpeg_encode): the encode MJPEGStreamer runs on
    the host (cv2.imencode, reference streamer.py:250-256), byte-identical to it for the same quality and restart interval.
    Owns its workspace and output buffer; `encode` is asynchronous on the current stream and returns (stream_buffer, size_tensor);
    `encode_bytes` synchronises and returns the JPEG as bytes."""

    def __init__(self, h: int, w: int, device, quality: int = 90, restart_interval: int = 4, capacity: int | None = None):
        L = _lib.lib()
        self.h, self.w, self.quality, self.restart_interval = int(h), int(w), int(quality), int(restart_interval)
        ws = L.d2s_jpeg_workspace_bytes(self.h, self.w, self.restart_interval)
        if ws == 0:
            raise ValueError(f"JpegEncoder: {L.d2s_last_error().decode('utf-8', 'replace')}")
        self.device = torch.device(device)
        self.capacity = int(capacity) if capacity else int(L.d2s_jpeg_max_bytes(self.h, self.w, self.restart_interval))
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
        self.out = torch.empty(self.capacity, dtype=torch.uint8, device=self.device)
        self.size = torch.zeros(1, dtype=torch.int32, device=self.device)

    def encode(self, rgb_hwc: torch.Tensor):
        _require_cuda(rgb_hwc, "rgb")
        if rgb_hwc.dtype != torch.uint8 or rgb_hwc.dim() != 3 or rgb_hwc.shape[2] != 3 or rgb_hwc.stride(2) != 1 or rgb_hwc.stride(1) != 3:
            raise ValueError("JpegEncoder.encode expects a uint8 [h,w,3] tensor with packed pixels")
        if tuple(rgb_hwc.shape[:2]) != (self.h, self.w):
            raise ValueError(f"JpegEncoder was built for {self.h}x{self.w}, got {tuple(rgb_hwc.shape[:2])}")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().d2s_jpeg_encode(rgb_hwc.data_ptr(), rgb_hwc.stride(0), self.h, self.w, self.quality, self.restart_interval,
                                                  self.out.data_ptr(), self.capacity, self.size.data_ptr(), self.workspace.data_ptr(),
                                                  self.workspace.numel(), _stream_ptr(self.device)), "d2s_jpeg_encode")
        return self.out, self.size

    def encode_bytes(self, rgb_hwc: torch.Tensor) -> bytes:
        out, size = self.encode(rgb_hwc)
        n = int(size.item())
        if n == 0:
            raise RuntimeError(f"JpegEncoder: the stream did not fit the {self.capacity}-byte buffer")
        return out[:n].cpu().numpy().tobytes()


def make_sbs_dibr(rgb: torch.Tensor, depth: torch.Tensor, ipd_uv=0.064, depth_ratio=1.0, convergence=0.0, display_mode="Half-SBS", *,
                  roll=0.0, resolution=None, search_radius=12, depth_tolerance=0.012, blur_radius=2.5, feather_enabled=False,
                  feather_width=0.0, corner_radius=0.0, rgb_layout="CHW", out: torch.Tensor | None = None, out_dtype=torch.float32,
                  out_layout="CHW", two_pass=True):
    """The reference viewer's occlusion-aware stereo rendering (viewer.py:386-631: disocclusion confidence + push-pull inpaint) as
    a tensor function: rgb [3,h,w] (0..255) + depth [h,w] -> packed frame [3,oh,ow] (or HWC), both eye views side by side / top-bottom
    as StereoWindow.render lays them out (viewer.py:2680-2760).  Defaults are the viewer's (ipd 0.064, depth_ratio 1.0, viewer.py:1326).
    `resolution` = u_resolution (default: the eye view's size)."""
    _require_cuda(rgb, "rgb"), _require_cuda(depth, "depth")
    if display_mode not in DISPLAY_MODES:
        raise ValueError(f"display_mode {display_mode!r}")
    h, w = rgb.shape[1:] if rgb_layout == "CHW" else rgb.shape[:2]
    depth = depth.squeeze().contiguous()
    if depth.dtype not in (torch.float32, torch.float16):
        depth = depth.float()
    if tuple(depth.shape) != (h, w):
        raise ValueError(f"depth must be [{h},{w}], got {tuple(depth.shape)}")
    L = _lib.lib()
    g = [C.c_int() for _ in range(4)]
    _lib.check(L.d2s_dibr_out_shape(h, w, DISPLAY_MODES[display_mode], *[C.byref(x) for x in g]), "d2s_dibr_out_shape")
    vh, vw, oh, ow = [x.value for x in g]
    if out is None:
        out = torch.empty((3, oh, ow) if out_layout == "CHW" else (oh, ow, 3), dtype=out_dtype, device=rgb.device)
    p = _lib.DibrParams()
    p.rgb, p.out = image_view(rgb, rgb_layout), image_view(out, out_layout)
    p.depth, p.depth_dtype, p.h, p.w = depth.data_ptr(), _TORCH2D2S[depth.dtype], h, w
    p.display_mode = DISPLAY_MODES[display_mode]
    p.ipd_uv, p.depth_ratio, p.convergence, p.roll = float(ipd_uv), float(depth_ratio), float(convergence), float(roll)
    p.resolution_x, p.resolution_y = (resolution if resolution is not None else (0.0, 0.0))
    p.search_radius, p.depth_tolerance, p.blur_radius = int(search_radius), float(depth_tolerance), float(blur_radius)
    p.feather_enabled, p.feather_width, p.corner_radius = int(bool(feather_enabled)), float(feather_width), float(corner_radius)
    if two_pass:      # scratch for the dense second pass over the pixels that need the inpaint sweep (per stream, like the other workspaces)
        from .prepost import _workspace
        ws = _workspace(rgb.device, L.d2s_dibr_workspace_bytes(h, w, DISPLAY_MODES[display_mode]), "dibr")
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(rgb.device):
        _lib.check(L.d2s_make_sbs_dibr(C.byref(p), _stream_ptr(rgb.device)), "d2s_make_sbs_dibr")
    return out


class _PinnedRing:
    """Ring of pinned host buffers for the float32 HWC result of make_sbs (depth.py:767-773).
    A frame handed out stays valid for `depth-1` further calls."""

    def __init__(self, depth=4):
        self.depth, self.bufs, self.i = depth, {}, 0

    def get(self, shape, dtype):
        key = (tuple(shape), dtype)
        ring = self.bufs.get(key)
        if ring is None:
            ring = self.bufs[key] = [torch.empty(shape, dtype=dtype, pin_memory=True) for _ in range(self.depth)]
        self.i = (self.i + 1) % self.depth
        return ring[self.i]


_ring = _PinnedRing()
_default_device = None


def default_device():
    global _default_device
    if _default_device is None:
        if not torch.cuda.is_available():
            raise _lib.D2SError("no CUDA device: desktop2stereo_b200 has no CPU path")
        _default_device = torch.device("cuda", torch.cuda.current_device())
    return _default_device


def make_sbs(rgb_c, depth, ipd_uv=0.064, depth_ratio=2.0, convergence=0.0, fill_16_9=False,
             display_mode="Half-SBS", fps=None, *, out_dtype=torch.float32):
    """depth.py:2186-2231.  rgb_c: np.ndarray [h,w,3] (RGB) or torch.Tensor [3,h,w]; depth: tensor or
    ndarray [h,w].  Returns the host float32 HWC frame the reference returns (`out_dtype=torch.uint8`
    gives the 12x smaller u8 frame instead)."""
    if isinstance(depth, np.ndarray):
        depth = torch.from_numpy(depth)
    dev = depth.device if depth.is_cuda else default_device()
    depth = depth.to(dev, non_blocking=True)
    if isinstance(rgb_c, np.ndarray):
        rgb = torch.from_numpy(rgb_c).to(dev, non_blocking=True)
        layout = "HWC" if (rgb.dim() == 3 and rgb.shape[2] == 3) else "CHW"
    else:
        rgb, layout = rgb_c.to(dev, non_blocking=True), "CHW"
    if fps is not None:
        from .overlay import overlay_fps
        rgb = overlay_fps(rgb, fps, layout=layout)
    # the reference casts rgb to depth.dtype before the warp (depth.py:2209-2215)
    round_rgb = rgb.dtype.is_floating_point and rgb.dtype != depth.dtype
    sbs = make_sbs_core(rgb, depth, ipd_uv=ipd_uv, depth_ratio=depth_ratio, display_mode=display_mode,
                        fill_16_9=fill_16_9, convergence=convergence, rgb_layout=layout,
                        out_dtype=out_dtype, out_layout="HWC", round_rgb_to_depth_dtype=round_rgb)
    host = _ring.get(sbs.shape, sbs.dtype)
    host.copy_(sbs, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host.numpy()
