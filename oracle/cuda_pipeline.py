"""ORACLE / BASELINE (test + bench infrastructure, never on the product path) — the reference's per-frame path as it runs on an
NVIDIA GPU, restated end to end with the reference's own library calls: the branches depth.py takes when IS_CUDA is true.

    process          depth.py:542-566    H2D of the captured frame, BGRA->RGB CHW, DTYPE (fp16)
    predict_depth    depth.py:1897-2025  bicubic-antialias resize (:698-699), /255, mean/std, model under torch.autocast("cuda")
                                         (:1763-1781 -> fp16 GEMMs/convs/SDPA via cuBLAS/cuDNN/flash), post_process_depth (:806-814),
                                         DepthStabilizer (:1865-1887), bilinear upsample (:1998-2004)
    make_sbs         depth.py:2186-2231  grid_sample warp under autocast (:2152-2160), cat, area pool, clamp,
                                         .cpu().float().permute(1,2,0).numpy() (:767-773)
The network is HF transformers' DepthAnythingForDepthEstimation itself (the reference's third-party dependency); pre/post/warp are
oracle/prepost.py and oracle/warp.py::make_sbs_core_torch, which are pinned on the unmodified reference (tests/golden).  Used by
`bench.py --impl reference-cuda` and the `reference_cuda` leg of the default line: "the bar to beat on B200 is the reference's own
torch CUDA path run on the same box" (SURVEY §2.2).  One frame at a time, like main.py's loop.
"""
from __future__ import annotations

import numpy as np
import torch

from . import prepost as opp
from .warp import make_sbs_core_torch


class ReferenceCUDAPipeline:
    def __init__(self, hf_model, device, depth_resolution=518, foreground_scale=0.05, aa_strength=4.0, dtype=torch.float16):
        self.dev = torch.device(device)
        self.model = hf_model.eval().float().to(self.dev)
        self.depth_resolution, self.fg, self.aa, self.dtype = depth_resolution, foreground_scale, aa_strength, dtype
        self.prev = None

    def process(self, frame_bgra: np.ndarray, target_height: int) -> torch.Tensor:
        t = torch.from_numpy(frame_bgra).to(self.dev, non_blocking=True)                  # depth.py:547
        return opp.process_cuda_branch(t, target_height, self.dtype)

    @torch.no_grad()
    def predict_depth(self, rgb: torch.Tensor, use_temporal_smooth=True) -> torch.Tensor:
        h, w = rgb.shape[1:]
        t = opp.resize_patch_aligned(rgb[None], self.depth_resolution, 14)                # fp16 in -> fp16 bicubic-antialias
        t = opp.normalise_input(t, torch.float32)
        with torch.autocast("cuda", enabled=True):                                        # DepthModelWrapper.__call__, depth.py:1763-1781
            depth = self.model(pixel_values=t).predicted_depth                            # fp16 out
        depth = opp.post_process_depth(depth, self.fg, self.aa)
        if use_temporal_smooth:
            self.prev, depth = opp.ema(self.prev, depth)
        return opp.upsample_depth(depth, h, w)

    @torch.no_grad()
    def make_sbs(self, rgb: torch.Tensor, depth: torch.Tensor, ipd_uv=0.064, depth_ratio=2.0, convergence=0.0, fill_16_9=False,
                 display_mode="Half-SBS") -> np.ndarray:
        with torch.autocast("cuda", enabled=True):
            out = make_sbs_core_torch(rgb.to(depth.dtype), depth, ipd_uv, depth_ratio, display_mode, fill_16_9, convergence)
        return out.cpu().float().permute(1, 2, 0).numpy()                    # chw_tensor_to_numpy, depth.py:767-773

    def frame(self, frame_bgra: np.ndarray, display_mode="Full-SBS", depth_ratio=2.0, **kw) -> np.ndarray:
        rgb = self.process(frame_bgra, frame_bgra.shape[0])
        depth = self.predict_depth(rgb, **kw)
        return self.make_sbs(rgb, depth, depth_ratio=depth_ratio, display_mode=display_mode)
