"""Drop-in mirror of the reference's `depth.py` hot-path interface, backed by libd2s_b200.

Public names and signatures follow the reference:
    process(img, target_height)                                   depth.py:542
    predict_depth(image_rgb, return_tuple=False, use_temporal_smooth=True, dtype=DTYPE)   depth.py:1897
    make_sbs_core(...), make_sbs(...)                              depth.py:2122 / 2186
    DepthModelWrapper / model_wraper                               depth.py:1539 / 1784  (engine slot)
so `main.py`'s loop (`process` -> `predict_depth` -> `make_sbs`, main.py:244-249, 1340) runs unchanged on top of it.
Two ways to use it:
    import desktop2stereo_b200.depth as depth; depth.init(hf_model)        # stand-alone module with the same names
    desktop2stereo_b200.depth.install(reference_depth_module, hf_model)    # patch the reference module in place
There is no CPU / PyTorch fallback: every function raises if the CUDA library or a GPU is missing.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .engine import B200Engine
from .prepost import IMAGENET_MEAN, IMAGENET_STD, PostProcessor, preprocess
from .prepost import process as _process
from .stereo import default_device, make_sbs, make_sbs_core  # noqa: F401  (re-exported, same names as the reference)


@dataclass
class Settings:
    """The constants the reference reads from settings.yaml through utils.py:819-907."""
    depth_resolution: int = 518        # DEPTH_RESOLUTION
    fp16: bool = True                  # FP16 -> DTYPE
    foreground_scale: float = 0.05     # settings["Foreground Scale"] / 10
    aa_strength: float = 4.0           # settings["Anti-aliasing"] * 2
    patch: int = 14                    # get_patch_size() for DA-V2 ids
    metric: bool = False


settings = Settings()
DTYPE = torch.float16
DEVICE = None


class DepthModelWrapper:
    """Engine slot with the reference wrapper's calling convention (depth.py:1763-1781): `__call__(tensor) -> depth`."""

    def __init__(self, engine: B200Engine):
        self.model = engine
        self.backend = B200Engine.backend_name
        self.device = engine.device
        self.dtype = DTYPE

    def __call__(self, tensor: torch.Tensor) -> torch.Tensor:
        return self.model(tensor)


model_wraper: DepthModelWrapper | None = None
depth_stabilizer: PostProcessor | None = None


def init(hf_model=None, *, engine: B200Engine | None = None, device=None, **overrides) -> DepthModelWrapper:
    """Build the engine from a transformers DepthAnythingForDepthEstimation (weights packed on the host) and
    set the module-level singletons the reference creates at import (depth.py:1784, 1890)."""
    global model_wraper, depth_stabilizer, DEVICE, DTYPE
    for k, v in overrides.items():
        if not hasattr(settings, k):
            raise TypeError(f"unknown setting {k!r}")
        setattr(settings, k, v)
    DTYPE = torch.float16 if settings.fp16 else torch.float32
    DEVICE = torch.device(device) if device is not None else default_device()
    if engine is None:
        if hf_model is None:
            raise ValueError("init() needs an HF model or a B200Engine")
        # the reference's CUDA path returns fp16 depth regardless of the FP16 flag (autocast)
        engine = B200Engine.from_hf_model(hf_model, DEVICE, out_dtype=torch.float16)
    model_wraper = DepthModelWrapper(engine)
    depth_stabilizer = PostProcessor(foreground_scale=settings.foreground_scale, aa_strength=settings.aa_strength,
                                     metric=settings.metric)
    return model_wraper


def _need_init():
    if model_wraper is None:
        raise _lib.D2SError("desktop2stereo_b200.depth.init(hf_model) has not been called")


def process(img, target_height: int) -> torch.Tensor:
    """depth.py:542-566 — BGRA/BGR u8 HWC frame -> RGB CHW tensor of DTYPE on the GPU (downscaled when target_height < h)."""
    return _process(img, target_height, dtype=DTYPE, device=DEVICE)


def predict_depth(image_rgb, return_tuple=False, use_temporal_smooth: bool = True, dtype=None, _backend_retry: bool = False):
    """depth.py:1897-2025 — returns depth in [0,1] at the frame's resolution ([h,w]), optionally with the RGB tensor."""
    _need_init()
    dev = model_wraper.device
    if isinstance(image_rgb, torch.Tensor):
        rgb_tensor = image_rgb.to(dev, non_blocking=True)
        h, w = rgb_tensor.shape[1:]
        layout = "CHW"
        src = rgb_tensor
    else:
        h, w = image_rgb.shape[:2]
        src = torch.from_numpy(image_rgb).to(dev, non_blocking=True)
        rgb_tensor = src.permute(2, 0, 1)
        layout = "HWC"
    x = preprocess(src, settings.depth_resolution, settings.patch, dtype=torch.float32, layout=layout,
                   mean=IMAGENET_MEAN, std=IMAGENET_STD)
    raw = model_wraper(x)                                  # [1,H',W'] fp16
    depth = depth_stabilizer(raw.reshape(raw.shape[-2:]), out_size=(h, w), use_temporal_smooth=use_temporal_smooth)
    if return_tuple:
        return depth, rgb_tensor
    return depth


_REF_SETTINGS = {"depth_resolution": "DEPTH_RESOLUTION", "fp16": "FP16", "foreground_scale": "FOREGROUND_SCALE", "aa_strength": "AA_STRENGTH"}
_PATCHED = ("process", "predict_depth", "make_sbs", "make_sbs_core")


def install(ref_depth_module, hf_model=None, *, engine: B200Engine | None = None, device=None, **overrides):
    """Patch an imported reference `depth` module so that main.py keeps calling its own names but lands on the B200 path:
    the engine goes into the wrapper slot exactly like TensorRTEngine would (depth.py:1597-1631), and the module-level
    `process`, `predict_depth`, `make_sbs`, `make_sbs_core` are replaced by the functions above.

    Call it once the reference module has executed COMPLETELY — as the last statement of depth.py (after make_sbs, :2231) or
    from main.py right after `import depth` and before `from depth import process, predict_depth` (main.py:44): the reference
    defines predict_depth (:1897), make_sbs_core (:2122) and make_sbs (:2186) below the place where it creates `model_wraper`
    (:1784), so a call placed there would be overwritten by those definitions.
    Settings the reference module holds as constants (DEPTH_RESOLUTION, FP16, FOREGROUND_SCALE, AA_STRENGTH; utils.py:819-907)
    are taken from it unless overridden."""
    missing = [n for n in _PATCHED if not hasattr(ref_depth_module, n)]
    if missing:
        raise _lib.D2SError(f"install(): {ref_depth_module.__name__} does not define {missing} (yet): call install() after the "
                            "reference module has executed completely")
    for key, const in _REF_SETTINGS.items():
        if key not in overrides and hasattr(ref_depth_module, const):
            overrides[key] = getattr(ref_depth_module, const)
    if hf_model is None and engine is None:
        hf_model = ref_depth_module.model_wraper.model
    w = init(hf_model, engine=engine, device=device, **overrides)
    ref_depth_module.model_wraper.model = w.model
    ref_depth_module.model_wraper.backend = w.backend
    for name in _PATCHED:
        setattr(ref_depth_module, name, globals()[name])
    return w
