"""Host-side mirror of the reference's overlay_fps (depth.py:2056-2103), backed by d2s_overlay_fps.

The reference keeps a module-level cache: the text mask is rebuilt on the first call and then only when its call counter is
a multiple of 10, so the number shown lags the fps argument by up to 9 frames.  The same state machine lives here; the
mask itself is never materialised (the kernel draws the glyph rectangle straight into the image)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_FPS_MASK_CACHE = {"text": None, "frame": 0, "interval": 10}


def reset_cache():
    _FPS_MASK_CACHE.update(text=None, frame=0)


def next_text(fps: float) -> bytes:
    """One step of the reference's text cache (depth.py:2061-2072): counts the call, refreshes the text on the first call and on
    every 10th; returns the text to draw (ASCII, <= 32 characters)."""
    cache = _FPS_MASK_CACHE
    cache["frame"] += 1
    if cache["text"] is None or cache["frame"] % cache["interval"] == 0:
        cache["text"] = f"FPS: {fps:.1f}"
    return cache["text"].encode("ascii", "replace")[:32]


def overlay_fps(rgb: torch.Tensor, fps: float, *, layout: str = "CHW", inplace: bool = False) -> torch.Tensor:
    """rgb [3,h,w] (or [h,w,3] with layout='HWC') on the GPU, any of f32/f16/bf16/u8.  Returns a new tensor like the
    reference does unless inplace=True (the pipeline owns its frame and skips the copy)."""
    from .stereo import _stream_ptr, image_view
    if not rgb.is_cuda:
        raise _lib.D2SError("overlay_fps: rgb must live on a CUDA device; there is no CPU path")
    cache = _FPS_MASK_CACHE
    next_text(fps)
    out = rgb if inplace else rgb.clone()
    h, w = (out.shape[1:] if layout == "CHW" else out.shape[:2])
    img = image_view(out, layout)
    text = cache["text"].encode("ascii", "replace")[:32]
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().d2s_overlay_fps(C.byref(img), int(h), int(w), text, _stream_ptr(out.device)), "d2s_overlay_fps")
    return out
