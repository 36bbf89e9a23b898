"""B200Engine — the object that plugs into the reference's engine slot.

The reference calls `model_wraper.model(tensor)` for every non-"PyTorch" backend (depth.py:1763-1781);
`TensorRTEngine.__call__` (depth.py:1457-1536) is the shape template: borrow `tensor.data_ptr()`, allocate the
output with torch on the same device, run on torch's current stream, return a tensor, offer `close()`.
This class does exactly that on top of `d2s_create / d2s_infer / d2s_destroy`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ModelConfig
from .stereo import _TORCH2D2S, _require_cuda, _stream_ptr, default_device
from .weights import config_from_hf, pack_state_dict


class B200Engine:
    backend_name = "B200"

    def __init__(self, blob: np.ndarray, cfg: ModelConfig, device=None, out_dtype=torch.float16):
        self.device = torch.device(device) if device is not None else default_device()
        if self.device.type != "cuda":
            raise _lib.D2SError("B200Engine needs a CUDA device; there is no CPU path")
        self.cfg, self.out_dtype = cfg, out_dtype
        self.policy = "latency"
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self._h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().d2s_create(blob.ctypes.data, blob.nbytes, C.byref(cfg), idx, C.byref(self._h)), "d2s_create")

    @classmethod
    def from_hf_model(cls, model, device=None, out_dtype=torch.float16):
        """model: transformers DepthAnythingForDepthEstimation (any device/dtype).  Weights are packed on the host."""
        cfg = config_from_hf(model.config)
        return cls(pack_state_dict(model.state_dict(), cfg), cfg, device, out_dtype)

    @classmethod
    def from_vda_state_dict(cls, state_dict, encoder: str, device=None, out_dtype=torch.float16):
        """state_dict of the reference's VideoDepthAnything (vda2_s.py; depth.py:870-902), encoder in vits/vitb/vitl.
        The engine is then a streaming one: one frame per call, state per CUDA stream, `reset()` starts a new video."""
        from .weights import config_for_vda, pack_vda_state_dict
        cfg = config_for_vda(encoder)
        return cls(pack_vda_state_dict(state_dict, cfg), cfg, device, out_dtype)

    @classmethod
    def from_state_dict(cls, state_dict, hf_config, device=None, out_dtype=torch.float16):
        cfg = config_from_hf(hf_config)
        return cls(pack_state_dict(state_dict, cfg), cfg, device, out_dtype)

    def __call__(self, tensor: torch.Tensor, out_dtype=None) -> torch.Tensor:
        """pixel_values [B,3,H,W] (fp16/fp32, normalised, H and W multiples of 14) -> predicted_depth [B,H,W]."""
        if self._h is None:
            raise _lib.D2SError("engine is closed")
        _require_cuda(tensor, "pixel_values")
        if tensor.dim() != 4 or tensor.shape[1] != 3:
            raise ValueError(f"pixel_values must be [B,3,H,W], got {tuple(tensor.shape)}")
        if tensor.dtype not in (torch.float16, torch.float32):
            tensor = tensor.float()
        tensor = tensor.contiguous()
        B, _, H, W = tensor.shape
        odt = out_dtype or self.out_dtype
        out = torch.empty((B, H, W), dtype=odt, device=tensor.device)
        with torch.cuda.device(tensor.device):
            _lib.check(_lib.lib().d2s_infer(self._h, tensor.data_ptr(), _TORCH2D2S[tensor.dtype], out.data_ptr(),
                                            _TORCH2D2S[odt], B, H, W, _stream_ptr(tensor.device)), "d2s_infer")
        return out.view(B, 1, H, W) if self.cfg.temporal else out     # VDA returns [T,1,H,W] (vda2_s.py:80-84)

    def set_policy(self, policy: str):
        """'latency' (default: one frame alone on the GPU) or 'throughput' (several frames in flight): tile shapes of the plans
        built by the next calls.  StereoPipeline switches to 'throughput' when it keeps more than one frame in flight."""
        _lib.check(_lib.lib().d2s_set_policy(self._h, {"latency": 0, "throughput": 1}[policy]), "d2s_set_policy")
        self.policy = policy

    def reset(self):
        """Temporal engines: forget the state of the current stream (the next frame is a first frame, vda2_s.py:196)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().d2s_reset_stream(self._h, _stream_ptr(self.device)), "d2s_reset_stream")

    def release_stream(self, stream=None):
        """Free the plans (and, for temporal engines, the video state) the engine keeps for a CUDA stream — call it before the
        stream is destroyed.  Host-synchronous."""
        ptr = (stream.cuda_stream if stream is not None else _stream_ptr(self.device))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().d2s_release_stream(self._h, ptr), "d2s_release_stream")

    def tap(self, name: str) -> torch.Tensor:
        """Debug/parity tap: an internal activation of the last inference, as a flat fp32 tensor."""
        n = C.c_size_t()
        L = _lib.lib()
        _lib.check(L.d2s_debug_tap(self._h, name.encode(), None, 0, C.byref(n), None), "d2s_debug_tap")
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.d2s_debug_tap(self._h, name.encode(), out.data_ptr(), n.value, C.byref(n),
                                       _stream_ptr(self.device)), "d2s_debug_tap")
        return out

    def workspace_bytes(self) -> int:
        return int(_lib.lib().d2s_workspace_bytes(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().d2s_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
