"""StereoPipeline — the caller side of the hot path: `depth` frames in flight over `depth` CUDA streams.

The reference's main.py runs capture -> depth -> warp as three threads joined by size-1 queues (main.py:67-68, 232-262,
1336-1341), i.e. a 3-deep software pipeline.  This class is the same idea on one GPU: each in-flight frame owns a CUDA
stream, a pinned host staging buffer for the captured BGRA frame, a pinned host buffer for the SBS result and (inside the
engine) its own activation buffers + CUDA graph, so H2D copy, network, warp and D2H copy of consecutive frames overlap.
Per frame it issues exactly the reference-facing calls: process -> predict_depth -> make_sbs_core (+ the host copy that
make_sbs does).  The only cross-frame dependency, the EMA of DepthStabilizer, is ordered with an event (prepost.py).
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch

from . import depth as d2s_depth
from .stereo import make_sbs_core, sbs_out_shape


class _Slot:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device)
        self.h_in = None       # pinned BGRA frame
        self.h_out = None      # pinned SBS result
        self.done = torch.cuda.Event()
        self.busy = False
        self.dev_out = None


class StereoPipeline:
    def __init__(self, depth_slots: int = 3, display_mode="Full-SBS", ipd_uv=0.064, depth_ratio=2.0, convergence=0.0,
                 fill_16_9=False, use_temporal_smooth=True, out_dtype=torch.float32, device=None):
        """`desktop2stereo_b200.depth.init(...)` must have been called (the engine and EMA state live there)."""
        d2s_depth._need_init()
        self.device = d2s_depth.model_wraper.device if device is None else torch.device(device)
        self.slots = [_Slot(self.device) for _ in range(depth_slots)]
        self.params = dict(ipd_uv=ipd_uv, depth_ratio=depth_ratio, convergence=convergence, fill_16_9=fill_16_9,
                           display_mode=display_mode)
        self.use_temporal_smooth, self.out_dtype = use_temporal_smooth, out_dtype
        self.next = 0
        self.pending: deque = deque()
        self.trace = None      # set to a list to collect per-frame CUDA events (start, after process, after depth, after warp)

    # ---- submission ----
    def _acquire(self) -> _Slot:
        s = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if s.busy:
            raise RuntimeError("pipeline full: collect a result before submitting another frame")
        s.busy = True
        return s

    def _enqueue(self, s: _Slot, frame_dev: torch.Tensor, to_host: bool):
        h = frame_dev.shape[0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.trace is not None else None
        if ev: ev[0].record(s.stream)
        rgb = d2s_depth.process(frame_dev, h)
        if ev: ev[1].record(s.stream)
        engine = d2s_depth.model_wraper.model
        prev = getattr(engine, "policy", None)
        if prev is not None and len(self.slots) > 1:
            engine.set_policy("throughput")        # several frames share the GPU: fewer, wider GEMM tiles
        try:
            depth = d2s_depth.predict_depth(rgb, use_temporal_smooth=self.use_temporal_smooth)
        finally:
            if prev is not None and len(self.slots) > 1:
                engine.set_policy(prev)
        if ev: ev[2].record(s.stream)
        sbs = make_sbs_core(rgb, depth, out_layout="HWC", out_dtype=self.out_dtype, **self.params)
        if ev:
            ev[3].record(s.stream)
            self.trace.append(ev)
        if to_host:
            if s.h_out is None or s.h_out.shape != sbs.shape or s.h_out.dtype != sbs.dtype:
                s.h_out = torch.empty(sbs.shape, dtype=sbs.dtype, pin_memory=True)
            s.h_out.copy_(sbs, non_blocking=True)
        s.dev_out = sbs
        s.done.record(s.stream)

    def submit(self, frame_bgra: np.ndarray):
        """Host frame (BGRA/BGR u8 HWC ndarray) -> ticket.  Copies the frame into this slot's pinned buffer, then enqueues
        H2D + process + predict_depth + make_sbs + D2H on the slot's stream; returns immediately."""
        s = self._acquire()
        src = torch.from_numpy(frame_bgra)
        if s.h_in is None or s.h_in.shape != src.shape:
            s.h_in = torch.empty(src.shape, dtype=torch.uint8, pin_memory=True)
        s.h_in.copy_(src)                                   # capture buffer -> pinned staging (host memcpy)
        with torch.cuda.stream(s.stream):
            frame_dev = s.h_in.to(self.device, non_blocking=True)
            self._enqueue(s, frame_dev, to_host=True)
        self.pending.append(s)
        return s

    def submit_pinned(self, frame_pinned: torch.Tensor):
        """Same, for a frame that already sits in pinned host memory (zero host-side copy)."""
        s = self._acquire()
        with torch.cuda.stream(s.stream):
            frame_dev = frame_pinned.to(self.device, non_blocking=True)
            self._enqueue(s, frame_dev, to_host=True)
        self.pending.append(s)
        return s

    def submit_device(self, frame_dev: torch.Tensor):
        """Frame already resident in HBM; the result stays on the device."""
        s = self._acquire()
        s.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s.stream):
            self._enqueue(s, frame_dev, to_host=False)
        self.pending.append(s)
        return s

    # ---- collection ----
    def result(self, ticket: _Slot | None = None, host: bool = True):
        """Blocks until the oldest (or the given) frame is complete; returns the float32 HWC ndarray (a view of the slot's
        pinned buffer, valid until the slot is reused `depth_slots` submissions later) or the device tensor."""
        s = ticket if ticket is not None else self.pending[0]
        s.done.synchronize()
        self.pending.remove(s)
        s.busy = False
        return s.h_out.numpy() if host else s.dev_out

    def run(self, frames, host: bool = True):
        """Generator: keeps the pipeline full while iterating `frames`; yields results in order."""
        def submit(f):
            if not host:
                return self.submit_device(f)
            if isinstance(f, torch.Tensor) and f.is_pinned():
                return self.submit_pinned(f)          # already in pinned host memory: no staging copy on the host
            return self.submit(f)
        for f in frames:
            if len(self.pending) == len(self.slots):
                yield self.result(host=host)
            submit(f)
        while self.pending:
            yield self.result(host=host)

    def out_shape(self, h, w):
        oh, ow = sbs_out_shape(h, w, self.params["display_mode"], self.params["fill_16_9"])
        return oh, ow, 3

