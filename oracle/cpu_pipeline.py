"""ORACLE (test infrastructure) — the reference's per-frame path as it runs on a CPU-only host, restated end to end:
process -> predict_depth -> make_sbs with the branches depth.py takes when IS_CUDA is false.

Used only by bench.py's `cpu_baseline` leg and `--impl reference` arm (and tests).  The depth network is HF transformers'
DepthAnythingForDepthEstimation itself (the reference's own third-party dependency), run under bf16 autocast exactly as
DepthModelWrapper.__call__ does on CPU (depth.py:661-664, 1763-1781).
Pinned: tests/test_oracle_cpu_pipeline.py runs the UNMODIFIED reference module (oracle/ref_harness.py) and this port on the same
frames and weights for 3 consecutive frames (EMA state included), Full- and Half-SBS: the float32 frames are bit-identical.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import prepost as opp
from .warp import make_sbs_core_torch


class ReferenceCPUPipeline:
    def __init__(self, hf_model, depth_resolution=518, foreground_scale=0.05, aa_strength=4.0, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.model = hf_model.eval().float()
        self.depth_resolution, self.fg, self.aa = depth_resolution, foreground_scale, aa_strength
        self.prev = None
        self.mean = torch.tensor(opp.IMAGENET_MEAN).view(1, 3, 1, 1)
        self.std = torch.tensor(opp.IMAGENET_STD).view(1, 3, 1, 1)

    @staticmethod
    def process(frame_bgra: np.ndarray, target_height: int) -> np.ndarray:
        """depth.py:570-629, ndarray path: cv2.cvtColor(BGRA2RGB); no resize when target_height >= h."""
        rgb = np.ascontiguousarray(frame_bgra[..., 2::-1])
        assert target_height >= rgb.shape[0], "the bench never downsizes the frame"
        return rgb

    def _resize(self, t: torch.Tensor) -> torch.Tensor:
        """depth.py:676-706, non-CUDA branch: strided pre-decimation then bilinear to the patch-aligned size."""
        _, _, h, w = t.shape
        nh, nw = opp.model_input_shape(h, w, self.depth_resolution, 14)
        if (nh, nw) == (h, w):
            return t
        longest = max(h, w)
        stride = longest // (self.depth_resolution * 2)
        if stride > 1:
            t = t[:, :, ::stride, ::stride]
        return F.interpolate(t.to(torch.float32), size=(nh, nw), mode="bilinear", align_corners=False)

    @torch.no_grad()
    def predict_depth(self, image_rgb: np.ndarray, use_temporal_smooth=True) -> torch.Tensor:
        h, w = image_rgb.shape[:2]
        t = torch.from_numpy(image_rgb).permute(2, 0, 1).unsqueeze(0)
        t = self._resize(t).to(torch.float32) / 255.0
        t = (t - self.mean) / self.std
        with torch.autocast(device_type="cpu", enabled=True):
            depth = self.model(pixel_values=t).predicted_depth
        depth = opp.post_process_depth(depth, self.fg, self.aa)
        if use_temporal_smooth:
            self.prev, depth = opp.ema(self.prev, depth)
        return opp.upsample_depth(depth, h, w)

    @torch.no_grad()
    def make_sbs(self, rgb: np.ndarray, depth: torch.Tensor, ipd_uv=0.064, depth_ratio=2.0, convergence=0.0,
                 fill_16_9=False, display_mode="Half-SBS") -> np.ndarray:
        """depth.py:2186-2231: rgb is cast to depth.dtype, grid_sample runs in fp32 under autocast."""
        r = torch.from_numpy(rgb).to(depth.dtype).permute(2, 0, 1).contiguous()
        out = make_sbs_core_torch(r, depth, ipd_uv, depth_ratio, display_mode, fill_16_9, convergence)
        return out.float().permute(1, 2, 0).numpy()

    def frame(self, frame_bgra: np.ndarray, display_mode="Full-SBS", depth_ratio=2.0, **kw) -> np.ndarray:
        rgb = self.process(frame_bgra, frame_bgra.shape[0])
        depth = self.predict_depth(rgb, **kw)
        return self.make_sbs(rgb, depth, depth_ratio=depth_ratio, display_mode=display_mode)
