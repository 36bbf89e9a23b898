"""StereoPipeline — the caller side of the hot path: `depth_slots` frames in flight, one C call per frame.

The reference's main.py runs capture -> depth -> warp as three threads joined by size-1 queues (main.py:67-68, 232-262,
1336-1341), i.e. a 3-deep software pipeline that issues ~15 framework calls per frame.  This class is a thin host-side
handle on `d2s_pipe_*` (csrc/pipe.cu): each in-flight frame owns a slot (CUDA stream, fixed device buffers, pinned host
buffers, CUDA graphs of the frame's kernels), so H2D copy, network, warp and D2H copy of consecutive frames overlap and the
host does one ctypes call per frame.  Per frame it computes exactly what the reference-facing calls compute:
process -> predict_depth -> make_sbs (+ the host copy make_sbs does) — same kernels, bit-identical results.  The only
cross-frame dependency, the EMA of DepthStabilizer (depth.py:1865-1887), is ordered frame-to-frame inside the pipe.
"""
from __future__ import annotations

import ctypes as C
from collections import deque

import numpy as np
import torch

from . import _lib
from . import depth as d2s_depth
from ._lib import DISPLAY_MODES, PipeConfig
from .prepost import IMAGENET_MEAN, IMAGENET_STD
from .stereo import _TORCH2D2S, sbs_out_shape

_NP_OF = {torch.float32: np.float32, torch.uint8: np.uint8, torch.float16: np.float16}


class _Ticket:
    __slots__ = ("pipe", "slot", "keep")

    def __init__(self, pipe, slot, keep=None):
        self.pipe, self.slot, self.keep = pipe, slot, keep


class _Pipe:
    """One d2s_pipe: fixed frame geometry and I/O mode."""

    def __init__(self, owner: "StereoPipeline", h0: int, w0: int, ch: int, host_io: bool):
        L = _lib.lib()
        s = d2s_depth.settings
        engine = d2s_depth.model_wraper.model
        c = PipeConfig()
        c.frame_h, c.frame_w, c.channels, c.target_height = h0, w0, ch, owner.target_height or h0
        c.rgb_dtype = _TORCH2D2S[d2s_depth.DTYPE]
        c.depth_resolution, c.patch = s.depth_resolution, s.patch
        c.mean[:], c.std[:] = IMAGENET_MEAN, IMAGENET_STD
        post = d2s_depth.depth_stabilizer
        c.metric, c.percentile, c.subsample_cap, c.gamma = int(post.metric), post.percentile, post.subsample_cap, post.gamma
        c.foreground_scale, c.aa_strength = post.foreground_scale, post.aa_strength
        c.use_temporal_smooth, c.ema_alpha = int(owner.use_temporal_smooth), post.ema_alpha
        p = owner.params
        c.ipd_uv, c.depth_ratio, c.convergence = float(p["ipd_uv"]), float(p["depth_ratio"]), float(p["convergence"])
        c.display_mode, c.fill_16_9 = DISPLAY_MODES[p["display_mode"]], int(bool(p["fill_16_9"]))
        c.out_dtype, c.slots, c.host_io = _TORCH2D2S[owner.out_dtype], owner.n_slots, int(host_io)
        c.streams = B = owner.streams
        c.out_format = {"rgb": 0, "nv12": 1, "jpeg": 2}[owner.out_format]
        c.jpeg_quality, c.jpeg_restart_interval = int(owner.jpeg_quality), int(owner.jpeg_restart_interval)
        self.jpeg = owner.out_format == "jpeg"
        c.fps_overlay = int(owner.show_fps)
        self.handle = C.c_void_p()
        self.engine = engine                 # the pipe borrows the engine's plans: keep it alive for as long as the pipe lives
        self.device, self.host_io, self.L = owner.device, host_io, L
        with torch.cuda.device(self.device):
            _lib.check(L.d2s_pipe_create(engine._h, C.byref(c), C.byref(self.handle)), "d2s_pipe_create")
        g = [C.c_int() for _ in range(6)]
        fb, ob = C.c_size_t(), C.c_size_t()
        _lib.check(L.d2s_pipe_geometry(self.handle, *[C.byref(x) for x in g], C.byref(fb), C.byref(ob)), "d2s_pipe_geometry")
        self.h, self.w, self.Hm, self.Wm, self.oh, self.ow = [x.value for x in g]
        self.frame_shape = (h0, w0, ch)
        self.host_in, self.host_out, self.dev_out, self.dev_depth, self.streams = [], [], [], [], []
        odt = owner.out_dtype
        for i in range(owner.n_slots):
            ptr = [C.c_void_p() for _ in range(6)]
            _lib.check(L.d2s_pipe_slot_buffers(self.handle, i, *[C.byref(x) for x in ptr]), "d2s_pipe_slot_buffers")
            hin, hout, _din, dout, ddep, st = [x.value for x in ptr]
            lead = (B,) if B > 1 else ()       # several streams: one frame of each per submit, stacked on a leading axis
            # jpeg: the raw d2s_pipe_jpeg_frame bytes of every stream (u32 size, 12 reserved bytes, then the stream)
            oshape = (ob.value,) if self.jpeg else (self.oh * 3 // 2, self.ow) if c.out_format == 1 else (self.oh, self.ow, 3)
            if host_io:   # numpy views of the library's pinned buffers (no copy)
                self.host_in.append(np.ctypeslib.as_array((C.c_uint8 * (B * fb.value)).from_address(hin)).reshape(lead + (h0, w0, ch)))
                raw = np.ctypeslib.as_array((C.c_uint8 * (B * ob.value)).from_address(hout))
                self.host_out.append(raw.view(_NP_OF[odt]).reshape(lead + oshape))
            self.dev_out.append(_device_view(dout, lead + oshape, odt, self.device))
            self.dev_depth.append(_device_view(ddep, lead + (self.h, self.w), torch.float16, self.device))
            self.streams.append(st)

    def close(self):
        if self.handle:
            self.host_in, self.host_out, self.dev_out, self.dev_depth, self.streams = [], [], [], [], []
            self.L.d2s_pipe_destroy(self.handle)
            self.handle = None
            self.engine = None


class _DevMem:
    """__cuda_array_interface__ holder: lets torch view memory the library owns (a pipe slot's device buffers)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def _device_view(ptr, shape, dtype, device):
    typestr = {torch.float32: "<f4", torch.float16: "<f2", torch.uint8: "|u1"}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_DevMem(ptr, shape, typestr), device=device)


def _jpeg_streams(buf, streams):
    """d2s_pipe_jpeg_frame buffers -> the JPEG stream(s) they hold (views; u8)"""
    def one(b):
        head = b[:4].cpu().numpy() if isinstance(b, torch.Tensor) else b[:4]
        return b[16:16 + int(np.frombuffer(head.tobytes(), np.uint32)[0])]
    return [one(buf[i]) for i in range(streams)] if streams > 1 else one(buf)


class StereoPipeline:
    def __init__(self, depth_slots: int = 3, display_mode="Full-SBS", ipd_uv=0.064, depth_ratio=2.0, convergence=0.0,
                 fill_16_9=False, use_temporal_smooth=True, out_dtype=torch.float32, device=None, target_height=None, streams=1,
                 out_format="rgb", jpeg_quality=90, jpeg_restart_interval=4, show_fps=False):
        """`desktop2stereo_b200.depth.init(...)` must have been called (the engine and the post-process settings live there).
        A temporal (Video-Depth-Anything) engine needs depth_slots == 1: its frames are sequential (vda2_s.py:189-224).
        streams > 1: that many concurrent video streams share the pipeline; every submit takes one frame of each, stacked as
        [streams, h, w, ch] (the network runs them as one batch; each stream has its own DepthStabilizer state).
        out_format="nv12" (with out_dtype=torch.uint8): results are NV12 frames [oh * 3 // 2, ow] u8 — the colour-conversion and
        4:2:0 stages of the JPEG encoder the reference runs on the host (streamer.py:250-256), done on the device: 1.5 B/px.
        out_format="jpeg" (with out_dtype=torch.uint8): results are complete JPEG streams (1-D u8 arrays; a list of them when
        streams > 1), byte-identical to cv2.imencode(".jpg", bgr, [IMWRITE_JPEG_QUALITY, jpeg_quality, IMWRITE_JPEG_RST_INTERVAL,
        jpeg_restart_interval]) of the u8 frame — MJPEGStreamer's whole encoder loop (streamer.py:231-256) on the device; what
        crosses PCIe is the compressed stream.
        show_fps=True: submit*(…, fps=value) draws the reference's "FPS: xx.x" overlay onto the frame before the warp, with
        make_sbs's semantics (depth.py:2226-2227; the text refreshes on every 10th call, overlay.py)."""
        d2s_depth._need_init()
        engine = d2s_depth.model_wraper.model
        if getattr(engine.cfg, "temporal", 0) and depth_slots != 1:
            raise _lib.D2SError("a Video-Depth-Anything engine keeps one video's 32-frame window per CUDA stream: frames of one "
                                "video are sequential, use depth_slots=1 (and one StereoPipeline per video)")
        if display_mode not in DISPLAY_MODES:
            raise ValueError(f"display_mode {display_mode!r}")
        self.device = d2s_depth.model_wraper.device if device is None else torch.device(device)
        self.n_slots, self.streams = depth_slots, int(streams)
        if out_format not in ("rgb", "nv12", "jpeg") or (out_format != "rgb" and out_dtype != torch.uint8):
            raise ValueError("out_format is 'rgb', 'nv12' or 'jpeg' (nv12 and jpeg need out_dtype=torch.uint8)")
        self.out_format, self.jpeg_quality, self.jpeg_restart_interval = out_format, jpeg_quality, jpeg_restart_interval
        self.show_fps = bool(show_fps)
        self.params = dict(ipd_uv=ipd_uv, depth_ratio=depth_ratio, convergence=convergence, fill_16_9=fill_16_9,
                           display_mode=display_mode)
        self.use_temporal_smooth, self.out_dtype, self.target_height = use_temporal_smooth, out_dtype, target_height
        self._pipes: dict = {}
        self._busy = [None] * depth_slots      # slot -> the _Pipe whose frame occupies it
        self.next = 0
        self.pending: deque = deque()
        self.trace = None      # set to a list to collect per-frame stage times (ms): process | network+post | upsample+warp

    # ---- plumbing ----
    def _pipe(self, shape, host_io) -> _Pipe:
        key = (tuple(shape), bool(host_io))
        p = self._pipes.get(key)
        if p is None:
            if self.streams > 1:
                if len(shape) != 4 or shape[0] != self.streams:
                    raise ValueError(f"expected [streams={self.streams}, h, w, 3|4] frames, got shape {tuple(shape)}")
                shape = shape[1:]
            if len(shape) != 3 or shape[2] not in (3, 4):
                raise ValueError(f"expected a uint8 [h,w,3|4] frame, got shape {tuple(shape)}")
            p = self._pipes[key] = _Pipe(self, shape[0], shape[1], shape[2], host_io)
        return p

    def _acquire(self) -> int:
        i = self.next
        if self._busy[i] is not None:
            raise RuntimeError("pipeline full: collect a result before submitting another frame")
        self.next = (self.next + 1) % self.n_slots
        return i

    def _submit(self, pipe: _Pipe, slot: int, ptr, ready_stream, keep=None, fps=None):
        L = pipe.L
        if (self.trace is not None) != getattr(pipe, "_tracing", False):
            pipe._tracing = self.trace is not None
            _lib.check(L.d2s_pipe_set_trace(pipe.handle, int(pipe._tracing)), "d2s_pipe_set_trace")
        if self.show_fps:        # make_sbs(fps=...): fps None -> no overlay on this frame (depth.py:2226)
            from .overlay import next_text
            _lib.check(L.d2s_pipe_set_fps_text(pipe.handle, next_text(fps) if fps is not None else None), "d2s_pipe_set_fps_text")
        elif fps is not None:
            raise ValueError("fps= needs StereoPipeline(show_fps=True)")
        with torch.cuda.device(self.device):
            _lib.check(L.d2s_pipe_submit(pipe.handle, slot, ptr, ready_stream), "d2s_pipe_submit")
        self._busy[slot] = pipe
        t = _Ticket(pipe, slot, keep)
        self.pending.append(t)
        return t

    # ---- submission ----
    def submit(self, frame_bgra: np.ndarray, fps=None):
        """Host frame (BGRA/BGR u8 HWC ndarray) -> ticket.  Copies the frame into the slot's pinned buffer (the capture ->
        staging memcpy), then enqueues H2D + process + predict_depth + make_sbs + D2H on the slot's stream; returns immediately."""
        if frame_bgra.dtype != np.uint8:
            raise ValueError(f"expected a uint8 frame, got {frame_bgra.dtype}")
        pipe = self._pipe(frame_bgra.shape, True)
        slot = self._acquire()
        np.copyto(pipe.host_in[slot], frame_bgra)
        return self._submit(pipe, slot, None, None, fps=fps)

    def submit_pinned(self, frame_pinned: torch.Tensor, fps=None):
        """Same, for a frame that already sits in pinned host memory (no host-side copy).  The tensor must stay alive and
        unchanged until the frame's result has been collected."""
        if not frame_pinned.is_pinned() or frame_pinned.dtype != torch.uint8 or not frame_pinned.is_contiguous():
            raise ValueError("submit_pinned needs a contiguous uint8 tensor in pinned host memory")
        pipe = self._pipe(frame_pinned.shape, True)
        slot = self._acquire()
        return self._submit(pipe, slot, frame_pinned.data_ptr(), None, keep=frame_pinned, fps=fps)

    def submit_device(self, frame_dev: torch.Tensor, fps=None):
        """Frame already resident in HBM (produced on torch's current stream); the result stays on the device."""
        if not frame_dev.is_cuda or frame_dev.dtype != torch.uint8 or not frame_dev.is_contiguous():
            raise _lib.D2SError("submit_device needs a contiguous uint8 CUDA tensor (there is no CPU path)")
        pipe = self._pipe(frame_dev.shape, False)
        slot = self._acquire()
        cur = torch.cuda.current_stream(self.device)
        # the ticket keeps the tensor alive until the frame has been collected, so the caching allocator cannot hand its memory to
        # someone else while the slot's stream still reads it (record_stream() on a library-owned stream would leave the allocator
        # holding a stream handle that the pipe destroys)
        return self._submit(pipe, slot, frame_dev.data_ptr(), cur.cuda_stream, keep=frame_dev, fps=fps)

    # ---- collection ----
    def result(self, ticket: _Ticket | None = None, host: bool = True):
        """Blocks until the oldest (or the given) frame is complete; returns the HWC ndarray (a view of the slot's pinned
        buffer, valid until the slot is reused `depth_slots` submissions later) or the device tensor (same lifetime)."""
        t = ticket if ticket is not None else self.pending[0]
        pipe, slot = t.pipe, t.slot
        _lib.check(pipe.L.d2s_pipe_wait(pipe.handle, slot), "d2s_pipe_wait")
        self.pending.remove(t)
        self._busy[slot] = None
        t.keep = None
        if self.trace is not None and getattr(pipe, "_tracing", False):
            ms = (C.c_float * 3)()
            _lib.check(pipe.L.d2s_pipe_slot_times(pipe.handle, slot, ms), "d2s_pipe_slot_times")
            self.trace.append(tuple(ms))
        out = pipe.host_out[slot] if host and pipe.host_io else pipe.dev_out[slot]
        return _jpeg_streams(out, self.streams) if pipe.jpeg else out

    def depth_of(self, ticket: _Ticket) -> torch.Tensor:
        """The [h,w] fp16 depth map (what predict_depth returns) of a collected frame; valid until the slot is reused."""
        return ticket.pipe.dev_depth[ticket.slot]

    def run(self, frames, host: bool = True, fps=None):
        """Generator: keeps the pipeline full while iterating `frames`; yields results in order.  fps (show_fps=True): a number
        or a callable returning the current rate, passed to every submit like main.py passes `current_fps` to make_sbs."""
        def submit(f):
            rate = fps() if callable(fps) else fps
            if not host:
                return self.submit_device(f, fps=rate)
            if isinstance(f, torch.Tensor):
                if f.is_pinned():
                    return self.submit_pinned(f, fps=rate)      # already in pinned host memory: no staging copy on the host
                f = f.numpy()
            return self.submit(f, fps=rate)
        for f in frames:
            if len(self.pending) == self.n_slots:
                yield self.result(host=host)
            submit(f)
        while self.pending:
            yield self.result(host=host)

    def reset(self):
        """A new video: forget the EMA state (and a temporal engine's window)."""
        while self.pending:
            self.result()
        for p in self._pipes.values():
            _lib.check(p.L.d2s_pipe_reset(p.handle), "d2s_pipe_reset")

    def out_shape(self, h, w):
        oh, ow = sbs_out_shape(h, w, self.params["display_mode"], self.params["fill_16_9"])
        return oh, ow, 3

    def close(self):
        while self.pending:
            self.result()
        for p in self._pipes.values():
            p.close()
        self._pipes = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
