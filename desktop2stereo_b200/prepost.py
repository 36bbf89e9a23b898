"""Host-side wrappers of the pre-/post-processing entry points of libd2s_b200.

    process                 depth.py:542-566
    model_input_shape       depth.py:676-692
    preprocess              depth.py:676-706 + :1931 + :1946-1948
    PostProcessor           depth.py:806-867, :1865-1887 (DepthStabilizer), :1998-2004
PyTorch appears only as the owner of device memory and of the current stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Image, PostParams
from .stereo import _TORCH2D2S, _require_cuda, _stream_ptr, default_device, image_view

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def process(img, target_height: int, dtype=torch.float16, device=None) -> torch.Tensor:
    """BGRA/BGR u8 HWC frame (np.ndarray or tensor) -> RGB CHW tensor of `dtype` on the GPU.
    Downscales with bilinear+antialias to an even size when target_height < frame height."""
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(img)
    dev = img.device if img.is_cuda else (device or default_device())
    img = img.to(dev, non_blocking=True)
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] not in (3, 4):
        raise ValueError(f"process expects a uint8 [h,w,3|4] frame, got {tuple(img.shape)} {img.dtype}")
    img = img.contiguous()
    H0, W0, ch = img.shape
    if target_height >= H0:
        h, w = H0, W0
    else:
        h = (target_height // 2) * 2
        w = (int(W0 * target_height / H0) // 2) * 2
    out = torch.empty((3, h, w), dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().d2s_process(img.data_ptr(), H0, W0, ch, out.data_ptr(), _TORCH2D2S[dtype], h, w,
                                          _stream_ptr(dev)), "d2s_process")
    return out


def model_input_shape(h: int, w: int, target: int = 518, patch: int = 14):
    nh, nw = C.c_int(), C.c_int()
    _lib.check(_lib.lib().d2s_model_input_shape(h, w, target, patch, C.byref(nh), C.byref(nw)), "d2s_model_input_shape")
    return nh.value, nw.value


_ws_cache: dict = {}


def _workspace(dev, nbytes: int, tag: str) -> torch.Tensor:
    key = (dev, tag, _stream_ptr(dev))   # per stream: concurrent frames must not share scratch
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _ws_cache[key] = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    return ws


def preprocess(rgb: torch.Tensor, target: int = 518, patch: int = 14, dtype=torch.float32, layout="CHW",
               mean=IMAGENET_MEAN, std=IMAGENET_STD, out: torch.Tensor | None = None) -> torch.Tensor:
    """rgb image (CHW / HWC / BGRA view, u8 or float, 0..255) -> normalised model input [1,3,H',W']."""
    _require_cuda(rgb, "rgb")
    h, w = (rgb.shape[1], rgb.shape[2]) if layout == "CHW" else (rgb.shape[0], rgb.shape[1])
    nh, nw = model_input_shape(h, w, target, patch)
    dev = rgb.device
    if out is None:
        out = torch.empty((1, 3, nh, nw), dtype=dtype, device=dev)
    L = _lib.lib()
    ws = _workspace(dev, L.d2s_preprocess_workspace_bytes(h, w, nh, nw), "pre")
    src = image_view(rgb, layout)
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    with torch.cuda.device(dev):
        _lib.check(L.d2s_preprocess(C.byref(src), h, w, out.data_ptr(), _TORCH2D2S[out.dtype], nh, nw, m, s,
                                    ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "d2s_preprocess")
    return out


class PostProcessor:
    """post_process_depth + DepthStabilizer + final upsample for ONE stream (the EMA state lives here)."""

    def __init__(self, foreground_scale=0.05, aa_strength=4.0, gamma=1.45, percentile=2.0, subsample_cap=6144,
                 ema_alpha=0.9, metric=False):
        self.foreground_scale, self.aa_strength, self.gamma = foreground_scale, aa_strength, gamma
        self.percentile, self.subsample_cap, self.ema_alpha, self.metric = percentile, subsample_cap, ema_alpha, metric
        self.prev: torch.Tensor | None = None   # DepthStabilizer.prev
        self.prev_valid = False

    def reset(self):
        self.prev, self.prev_valid = None, False

    def __call__(self, depth: torch.Tensor, out_size=None, use_temporal_smooth=True, out_dtype=None,
                 return_lowres=False, compute_dtype=None):
        """depth: raw predicted_depth [H,W] (or [1,H,W]).  Returns the [h,w] map predict_depth would return
        (and the low-res post-processed map if asked)."""
        _require_cuda(depth, "depth")
        d = depth.squeeze()
        if d.dim() != 2:
            raise ValueError(f"depth must be [H,W], got {tuple(depth.shape)}")
        d = d.contiguous()
        H, W = d.shape
        dev = d.device
        cdt = compute_dtype or d.dtype
        odt = out_dtype or cdt
        oh, ow = out_size if out_size is not None else (H, W)
        L = _lib.lib()
        ws = _workspace(dev, L.d2s_postprocess_workspace_bytes(H, W), "post")
        out = torch.empty((oh, ow), dtype=odt, device=dev)
        low = torch.empty((H, W), dtype=cdt, device=dev) if return_lowres else None
        p = PostParams()
        p.depth_in, p.in_dtype, p.H, p.W = d.data_ptr(), _TORCH2D2S[d.dtype], H, W
        p.out, p.out_dtype, p.out_h, p.out_w = out.data_ptr(), _TORCH2D2S[odt], oh, ow
        p.compute_dtype, p.metric = _TORCH2D2S[cdt], int(self.metric)
        p.percentile, p.subsample_cap, p.gamma = self.percentile, self.subsample_cap, self.gamma
        p.foreground_scale, p.aa_strength = self.foreground_scale, self.aa_strength
        cur = torch.cuda.current_stream(dev)
        if use_temporal_smooth:
            if self.prev is None or self.prev.shape != (H, W) or self.prev.device != dev or self.prev.dtype != cdt:
                # (DepthStabilizer resets on a shape/device change, depth.py:1878-1882.)  Earlier frames may still be updating
                # the old state on other streams: let them finish before its memory goes back to the allocator
                if getattr(self, "_ema_event", None) is not None:
                    self._ema_event.synchronize()
                self.prev, self.prev_valid = torch.empty((H, W), dtype=cdt, device=dev), False
                self._ema_event = None
            p.ema_state, p.ema_valid, p.ema_alpha = self.prev.data_ptr(), int(self.prev_valid), self.ema_alpha
            self.prev_valid = True
            # frames of one video stream may be in flight on several CUDA streams: the EMA is the one cross-frame
            # dependency of the path (depth.py:1865-1887), so frame t's update is ordered after frame t-1's
            if getattr(self, "_ema_event", None) is not None:
                cur.wait_event(self._ema_event)
        p.out_lowres = low.data_ptr() if low is not None else None
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(dev):
            _lib.check(L.d2s_postprocess(C.byref(p), cur.cuda_stream), "d2s_postprocess")
        if use_temporal_smooth:
            self._ema_event = torch.cuda.Event()
            self._ema_event.record(cur)
        return (out, low) if return_lowres else out
