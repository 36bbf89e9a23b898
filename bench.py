#!/usr/bin/env python
"""bench.py — end-to-end frames/s of the depth-inference + SBS-warp hot path (BASELINE.json metric: "at 1080p & 4K").

Headline workload at every N (`value`, `e2e`): BASELINE.json configs[1] per GPU — Depth Anything V2-Base, 1080p BGRA frames,
batch 1, Full-SBS (model input 294x518, 778 tokens).  The 4K half of the metric rides on the same line as the `large4k` block:
configs[2] per GPU — DA-V2-Large, 8 concurrent 4K streams batched through the network (M = 6224 token rows), Full-SBS — with its own
`value`, `e2e`, `roofline`, `cpu_baseline`; under torchrun the `config5` block is BASELINE configs[4]: 8 streams x 8 frames of 4K,
stream s -> rank s mod G (strong scaling).  Synthetic frames, seeded random-init weights (no network).  Frames/streams shard across
ranks with no per-frame collective ("scaling": "weak"); the only collective is one NCCL broadcast of the packed weights at init.

    python bench.py --gpus 1 --steps 200 --warmup 10            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1        # the reference's CPU path (oracle port), host cores
    python bench.py --impl reference-cuda --steps 50             # the reference's torch-CUDA path (HF fp16 autocast + ATen) on this GPU
One JSON line on stdout (rank 0).  `value` = frames/s with frames resident in HBM; `e2e` = the same through the repo's public API
(StereoPipeline over d2s_pipe_*) with HOST buffers: H2D of the pinned BGRA frame and D2H of the float32 SBS frame inside the timed
region.  A timed region is EXACTLY --steps frames between barrier + synchronize on both sides (max over ranks); it is repeated
until >= 1 s of device time has been measured and the MEDIAN region is reported (`repeats`, `region_ms_min/max` give the spread).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 1080, 1920
VARIANT = "Base"
DISPLAY_MODE = "Full-SBS"
DEPTH_RATIO = 2.0
SEED = 0
METRIC = "end-to-end frames/sec (depth infer + SBS warp)"


def model_flops(cfg, Hm, Wm):
    """Algorithmic FLOPs (2*MAC) of one frame through the network (SURVEY §8d formulas)."""
    D, L, F = cfg.hidden, cfg.layers, cfg.fusion
    c = [cfg.neck[i] for i in range(4)]
    ph, pw = Hm // 14, Wm // 14
    P, N = ph * pw, ph * pw + 1
    fl = 2 * P * 588 * D + L * (24 * N * D * D + 4 * N * N * D)
    hs = [4 * ph, 2 * ph, ph, (ph - 1) // 2 + 1]
    ws = [4 * pw, 2 * pw, pw, (pw - 1) // 2 + 1]
    fl += sum(2 * P * D * ci for ci in c) + 2 * P * c[0] * 16 * c[0] + 2 * P * c[1] * 4 * c[1] + 2 * hs[3] * ws[3] * 9 * c[3] * c[3]
    fl += sum(2 * hs[i] * ws[i] * 9 * c[i] * F for i in range(4))
    for j in range(4):
        lv = 3 - j
        n_rcu = 1 if j == 0 else 2
        fl += n_rcu * 2 * (2 * hs[lv] * ws[lv] * 9 * F * F)
        oh, ow = (hs[lv - 1], ws[lv - 1]) if j < 3 else (2 * hs[lv], 2 * ws[lv])
        fl += 2 * oh * ow * F * F
    fl += 2 * (8 * ph) * (8 * pw) * 9 * F * (F // 2) + 2 * Hm * Wm * 9 * (F // 2) * 32 + 2 * Hm * Wm * 32
    return fl


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line): NVML polled at 10 Hz from a
    thread of this process (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints, without
    spawning a process per leg); `nvidia-smi -lms 100` is the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.proc, self.lines, self.nv, self.samples = index, None, [], None, []

    def _nvml_handle(self, nv):
        try:        # by PCI address, so a CUDA_VISIBLE_DEVICES remap cannot point at the wrong board
            import torch
            pr = torch.cuda.get_device_properties(self.index)
            return nv.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:      # noqa: BLE001
            return nv.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handle = self._nvml_handle(nv)
            nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
            self.nv, self._stop = nv, threading.Event()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:      # noqa: BLE001 - any NVML problem: the nvidia-smi path below
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.handle
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        while not self._stop.is_set():
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                                     [n for n, b in zip(self.NAMES, bits) if r & b]))
            except Exception:      # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        sm, mx, reasons = [], [], set()
        if self.nv is not None:
            time.sleep(0.15)
            self._stop.set()
            self.thread.join(timeout=2.0)
            for a, b, rs in self.samples:
                sm.append(float(a)); mx.append(float(b)); reasons.update(rs)
            source = "nvml"
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.15)
            self.proc.terminate()
            for ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(self.NAMES, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            source = "nvidia-smi"
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": source}



def build_hf_model(variant=VARIANT, seed=SEED):
    from desktop2stereo_b200.synth import make_hf_model   # seeded random-init weights (no checkpoints offline)
    return make_hf_model(variant, seed)


# ---------------------------------------------------------------------------------------------------------------------
# reference arms (oracle/: the checker and the baselines — never the product path)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_fps(n_frames, warmup, threads, variant=VARIANT, h=H, w=W, mode=DISPLAY_MODE):
    """The reference's CPU path (oracle port: HF model under bf16 autocast + torch pre/post/warp) on `threads` host cores."""
    import numpy as np
    import torch
    from oracle.cpu_pipeline import ReferenceCPUPipeline
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        pipe = ReferenceCPUPipeline(build_hf_model(variant), 518)
        rng = np.random.default_rng(SEED)
        frames = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for _ in range(2)]
        for i in range(warmup):
            pipe.frame(frames[i % 2], mode, DEPTH_RATIO)
        t0 = time.perf_counter()
        for i in range(n_frames):
            out = pipe.frame(frames[i % 2], mode, DEPTH_RATIO)
        dt = time.perf_counter() - t0
    finally:
        torch.set_num_threads(old)
    return n_frames / dt, dt


def cuda_reference_fps(n_frames, warmup, dev, variant=VARIANT, h=H, w=W, mode=DISPLAY_MODE):
    """The reference's torch-CUDA path restated with its own library calls (oracle/cuda_pipeline.py): HF module under fp16
    autocast (cuBLAS / cuDNN / SDPA), ATen resize / topk / conv2d / grid_sample, one frame at a time with the H2D of the frame and
    the .cpu() of the float32 result per frame — what main.py's loop does on an NVIDIA GPU."""
    import numpy as np
    import torch
    from oracle.cuda_pipeline import ReferenceCUDAPipeline
    pipe = ReferenceCUDAPipeline(build_hf_model(variant), dev, 518)
    rng = np.random.default_rng(SEED)
    frames = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for _ in range(4)]
    for i in range(warmup):
        pipe.frame(frames[i % 4], mode, DEPTH_RATIO)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(n_frames):
        out = pipe.frame(frames[i % 4], mode, DEPTH_RATIO)       # ends in .cpu(): host-synchronous, like make_sbs
    dt = time.perf_counter() - t0
    assert out.shape[2] == 3
    del pipe
    torch.cuda.empty_cache()
    return n_frames / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    warm = max(args.warmup, 1)
    if args.impl == "reference-cuda":
        import torch
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        fps, dt = cuda_reference_fps(args.steps, max(args.warmup, 10), dev)   # cuDNN autotuning and clock ramp-up settle within ~10 frames
        extra = {"impl": "reference-cuda", "dtype": "fp16 autocast (torch CUDA: cuBLAS/cuDNN/SDPA + ATen)",
                 "reference_cuda": {"value": fps, "unit": "frames/s", "kind": "port", "sample": f"{args.steps} frames, one at a time, H2D + .cpu() per frame"},
                 "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": H * W * 4, "d2h_bytes_per_step": H * 2 * W * 3 * 4}}
    else:
        fps, dt = cpu_reference_fps(args.steps, warm, threads)
        extra = {"impl": "reference", "dtype": "bf16 autocast (CPU)",
                 "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                  "sample": f"{args.steps} frames of the same 1080p workload, 1 frame per step"},
                 "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": WORKLOADS["base1080"][4] + ", per GPU"}}
    line.update(extra)
    print(json.dumps(line))
    return 0


WORKLOADS = {
    # name: (variant, frame h, frame w, streams per submit, description)
    "base1080": ("Base", 1080, 1920, 1, "BASELINE.json configs[1]: DA-V2-Base, 1080p BGRA batch=1 -> Full-SBS (model input 294x518, 778 tokens)"),
    "large4k": ("Large", 2160, 3840, 8, "BASELINE.json configs[2]: DA-V2-Large, 4K BGRA batch=8 -> Full-SBS (model input 8 x 294x518, 6224 token rows)"),
    "vda1080": ("vits", 1080, 1920, 1, "BASELINE.json configs[3]: streaming Video-Depth-Anything, 1080p, 32-frame temporal window, one video per CUDA stream"),
}


# ---------------------------------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------------------------------
def write_all(fd, text):
    data = text.encode()
    while data:
        data = data[os.write(fd, data):]


class Watchdog:
    """bench.py must end with its JSON line.  If no timed region completes for `limit` seconds (a whole default run takes ~90 s),
    something is stuck: in a block AFTER the headline has been measured, rank 0 prints the line it has — the stalled block named
    in "aborted" — and every rank leaves with status 0; before that, the ranks leave with status 3 instead of waiting ten
    minutes for the NCCL watchdog."""

    LIMITS = (240.0, 600.0)       # seconds without a completed timed region: in a block after the headline / before it
    POLL = 5.0

    def __init__(self):
        self.t = time.monotonic()
        self.phase, self.optional, self.finalize, self.fd, self.rank, self.done, self.disabled = "setup", False, None, 1, 0, False, False
        threading.Thread(target=self._run, daemon=True).start()

    def beat(self):
        self.t = time.monotonic()

    def enter(self, phase, optional):
        self.phase, self.optional = phase, optional
        self.beat()

    def _run(self):
        while True:
            time.sleep(self.POLL)
            limit = self.LIMITS[0] if self.optional else self.LIMITS[1]
            if self.disabled or time.monotonic() - self.t <= limit:
                continue
            msg = f"no timed region completed for {limit:g} s in block '{self.phase}'"
            os.write(2, f"[bench watchdog] rank {self.rank}: {msg}\n".encode())
            if self.done:
                os._exit(0)
            if self.optional and self.finalize is not None:
                if self.rank == 0:
                    try:
                        write_all(self.fd, json.dumps(self.finalize(aborted=msg)) + "\n")
                    except Exception as exc:      # noqa: BLE001 - nothing else can be done here
                        os.write(2, f"[bench watchdog] could not finalize the line: {exc!r}\n".encode())
                        os._exit(3)
                os._exit(0)
            os._exit(3)


WATCHDOG = None


class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args, self.torch, self.dist = args, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        # one process per GPU, pinned to the cores (and, by first touch, the memory) of the GPU's NUMA node — before any pinned
        # buffer exists
        from desktop2stereo_b200.sharding import bind_to_gpu_numa
        self.numa = bind_to_gpu_numa(self.local, self.local, int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world))))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.warmup = max(args.warmup, 3)
        self.peaks, self.peak_src = load_peaks()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def region(self, fn, steps):
        """EXACTLY `steps` steps through fn, device-timed between barrier + synchronize on both sides, max over ranks -> ms"""
        torch = self.torch
        self.barrier(); torch.cuda.synchronize()
        s, e = self.ev(), self.ev()
        s.record()
        fn(range(steps))
        torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize(); self.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        if WATCHDOG is not None:
            WATCHDOG.beat()
        return ms.item()

    def timed(self, fn, steps, min_s=1.0, max_repeats=400):
        """the K-step region repeated until >= min_s of device time: median region + spread (every rank runs the same count)"""
        first = self.region(fn, steps)
        reps = int(min(max(math.ceil(min_s * 1e3 / max(first, 1e-3)), 3), max_repeats))
        ms = sorted([first] + [self.region(fn, steps) for _ in range(reps - 1)])
        return {"ms": ms[len(ms) // 2], "min": ms[0], "max": ms[-1], "repeats": len(ms)}


def desktop_like_frames(torch, n, streams, h, w, seed):
    """BGRA frames with the statistics of desktop content (flat panels, gradients, 1-pixel text-like strokes, one noisy 'video'
    window): the size of a JPEG stream depends on content, so the compact-output legs are measured on this AND on uniform noise.
    Built at 1080p and tiled to larger sizes.  Pinned host tensors [streams, h, w, 4] (or [h, w, 4])."""
    import numpy as np
    bh, bw = min(h, 1080), min(w, 1920)

    def one(sd):
        rng = np.random.default_rng(sd)
        yy, xx = np.mgrid[0:bh, 0:bw].astype(np.float32)
        img = np.stack([40 + 60 * xx / bw, 50 + 80 * yy / bh, 90 + 40 * (xx + yy) / (bh + bw)], -1)
        for _ in range(24):
            y0, x0 = int(rng.integers(0, bh - 8)), int(rng.integers(0, bw - 8))
            img[y0:y0 + int(rng.integers(8, bh // 3)), x0:x0 + int(rng.integers(8, bw // 3))] = rng.integers(0, 256, 3)
        panel = img[bh // 8:bh // 8 + bh // 4, bw // 8:bw // 8 + bw // 3]
        panel[:] = 235
        mask = rng.random(panel.shape[:2]) < 0.18
        mask[::3] = False
        panel[mask] = 20
        vy, vx, vh, vw = bh // 2, bw // 2, bh // 3, bw // 3
        tex = 128 + 60 * np.sin(xx[vy:vy + vh, vx:vx + vw] / 9.0) * np.cos(yy[vy:vy + vh, vx:vx + vw] / 7.0)
        img[vy:vy + vh, vx:vx + vw] = tex[..., None] + rng.normal(0, 6, (vh, vw, 3))
        img = np.clip(img, 0, 255).astype(np.uint8)
        img = np.tile(img, (-(-h // bh), -(-w // bw), 1))[:h, :w]
        return np.concatenate([img[..., ::-1], np.full((h, w, 1), 255, np.uint8)], -1)
    base = [one(seed + i) for i in range(4)]
    out = []
    for i in range(n):
        f = np.stack([base[(i + b) % 4] for b in range(streams)]) if streams > 1 else base[i % 4]
        out.append(torch.from_numpy(np.ascontiguousarray(f)).pin_memory())
    return out


def jpeg_leg(d, **kw):
    keys = ("h2d_bytes_per_step", "d2h_bytes_per_step", "pcie_gbs_per_gpu", "jpeg_bytes_per_frame", "jpeg_bytes_per_pixel", "content", "jpeg")
    return dict({"value": d["fps"], "unit": "frames/s"}, **{k: d[k] for k in keys}, **kw)


def pipe_legs(B, depth, frames_dev, frames_host, h, w, streams, slots, steps, out_dtypes, want_device=True):
    """device-resident and host-buffer legs of one workload through StereoPipeline; a step = one submit = `streams` frames.
    Legs named e2e_jpeg*: out_format='jpeg' (the whole JPEG encode on the device; d2h bytes = the measured stream sizes, rounded up
    as the pipe's adaptive copy does); *_desktop: desktop-like frames instead of uniform noise."""
    import numpy as np
    torch = B.torch
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.pipeline import StereoPipeline
    L = _lib.lib()
    res = {}
    ring_d, ring_h = len(frames_dev), len(frames_host)
    warm = max(B.warmup, 2 * slots)
    fpp = streams                                  # frames per step
    oshape = ((streams,) if streams > 1 else ()) + (h, 2 * w, 3)
    if want_device:
        pipe = StereoPipeline(depth_slots=slots, display_mode=DISPLAY_MODE, depth_ratio=DEPTH_RATIO, streams=streams)

        def run_dev(idx):
            for _ in pipe.run((frames_dev[i % ring_d] for i in idx), host=False):
                pass
        run_dev(range(warm))
        torch.cuda.synchronize()
        clocks = ClockSampler(B.local); clocks.start()
        l0 = L.d2s_launch_count()
        pipe.trace = []
        torch.cuda.profiler.start()          # no-op unless run under `ncu --profile-from-start off`
        t = B.timed(run_dev, steps)
        torch.cuda.profiler.stop()
        launches_per_region = (L.d2s_launch_count() - l0) / t["repeats"]
        res["clocks"] = clocks.stop()
        st = np.array(pipe.trace) if pipe.trace else np.zeros((1, 3))
        pipe.trace = None
        res["device"] = dict(t, fps=B.world * steps * fpp / (t["ms"] / 1e3), launches=int(round(launches_per_region)),
                             stage_ms=dict(zip(["process", "resize+network+postprocess", "upsample+warp"], np.median(st, 0).tolist())))
        pipe.close()
    for name, odt in out_dtypes:
        fmt = "jpeg" if "jpeg" in name else "nv12" if name.endswith("nv12") else "rgb"
        src = desktop_like_frames(torch, ring_h, streams, h, w, SEED + 7 * B.rank) if name.endswith("_desktop") else frames_host
        if fmt == "nv12":
            oshape = ((streams,) if streams > 1 else ()) + (h * 3 // 2, 2 * w)
        pipe = StereoPipeline(depth_slots=slots, display_mode=DISPLAY_MODE, depth_ratio=DEPTH_RATIO, out_dtype=odt, streams=streams, out_format=fmt)
        sizes = []

        def run_host(idx, p=pipe, src=src, fmt=fmt):
            r = None
            for r in p.run((src[i % ring_h] for i in idx), host=True):   # pinned host frames -> H2D inside the timed region
                if fmt == "jpeg":
                    sizes.extend(len(x) for x in (r if streams > 1 else [r]))
            return r
        r = run_host(range(warm))
        if fmt == "jpeg":
            for x in (r if streams > 1 else [r]):
                assert bytes(x[:3]) == b"\xff\xd8\xff" and bytes(x[-2:]) == b"\xff\xd9"
        else:
            assert tuple(r.shape) == oshape and r.dtype == {torch.float32: np.float32, torch.uint8: np.uint8}[odt]
        sizes.clear()
        t = B.timed(run_host, steps)
        es = 4 if odt == torch.float32 else 1
        h2d = fpp * h * w * 4
        extra = {}
        if fmt == "jpeg":
            mean = float(np.mean(sizes))
            d2h = int(fpp * ((int(mean * 1.25) + 16 + 65535) // 65536) * 65536)
            extra = {"jpeg_bytes_per_frame": mean, "jpeg_bytes_per_pixel": mean / (h * 2 * w), "content": "desktop-like" if name.endswith("_desktop") else "uniform noise",
                     "jpeg": "quality 90, restart interval 4 MCUs; byte-identical to cv2.imencode (tests/test_jpeg_gpu.py)"}
        else:
            d2h = fpp * h * 2 * w * 3 * es if fmt == "rgb" else fpp * h * 2 * w * 3 // 2
        fps = B.world * steps * fpp / (t["ms"] / 1e3)
        res[name] = dict(t, fps=fps, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, pcie_gbs_per_gpu=(h2d + d2h) * fps / fpp / B.world / 1e9, **extra)
        pipe.close()
    return res


def host_copy_bandwidth(B, nbytes=256 << 20):
    """measured D2H / H2D rates of this rank's pinned memory — each direction alone and both at once, every rank copying at the same
    time (max over ranks of the elapsed time): names the limiter of the e2e rows.  `aggregate_both_gbs` = all ranks, both directions."""
    torch = B.torch
    hbuf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); hbuf2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dbuf = torch.empty(nbytes, dtype=torch.uint8, device=B.dev); dbuf2 = torch.empty(nbytes, dtype=torch.uint8, device=B.dev)
    s2 = torch.cuda.Stream(B.dev)
    out = {}
    for name, (dst, src) in {"d2h_gbs": (hbuf, dbuf), "h2d_gbs": (dbuf, hbuf)}.items():
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        ms = B.region(lambda idx: [dst.copy_(src, non_blocking=True) for _ in idx], 8)
        out[name] = 8 * nbytes / (ms / 1e3) / 1e9

    def both(idx):
        for _ in idx:
            hbuf.copy_(dbuf, non_blocking=True)
            with torch.cuda.stream(s2):
                dbuf2.copy_(hbuf2, non_blocking=True)
        torch.cuda.current_stream(B.dev).wait_stream(s2)
    both(range(1)); torch.cuda.synchronize()
    ms = B.region(both, 8)
    out["both_directions_gbs_per_gpu"] = 2 * 8 * nbytes / (ms / 1e3) / 1e9
    out["aggregate_both_gbs"] = out["both_directions_gbs_per_gpu"] * B.world
    out["note"] = "pinned-memory copies of 256 MB, all ranks at once; the e2e rows move h2d_bytes_per_step + d2h_bytes_per_step per frame through this"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large4k", action="store_true", help="skip the 4K (configs[2] / configs[4]) block")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=3)
    ap.add_argument("--slots", type=int, default=8, help="frames in flight (CUDA streams) in the pipelined legs")
    ap.add_argument("--vda-encoder", default="vits", choices=["vits", "vitb", "vitl"])
    ap.add_argument("--workload", default="base1080", choices=["base1080", "vda1080"],
                    help="base1080 is the headline (configs[1], with the large4k block); vda1080 (configs[3]) is an extra measurement")
    args = ap.parse_args()
    if args.impl != "b200" or args.workload == "vda1080":
        if WATCHDOG is not None:
            WATCHDOG.disabled = True          # the reference arm's steps are CPU work of unknown length; nothing there can stall on a GPU
        return run_reference(args) if args.impl != "b200" else run_vda1080(args)

    import numpy as np
    B = Bench(args)
    if WATCHDOG is not None:
        WATCHDOG.enter("headline (configs[1])", False)
    torch, dev, world, rank = B.torch, B.dev, B.world, B.rank
    from desktop2stereo_b200 import _lib, depth
    from desktop2stereo_b200.stereo import make_sbs_core
    L = _lib.lib()
    peaks, src = B.peaks, B.peak_src

    # ================= headline: configs[1] =================
    engine, cfg = build_engine(VARIANT, rank, world, dev)
    depth.init(engine=engine, device=dev)
    RING = 24          # a ring larger than L2 so no timed iteration re-reads a cached frame
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    frames = [torch.randint(0, 256, (H, W, 4), generator=g, dtype=torch.uint8, device=dev) for _ in range(RING)]
    host_frames = [f.cpu().pin_memory() for f in frames[:4]]   # pinned host copies for the end-to-end legs
    host_np = [t.numpy() for t in host_frames]

    # ---- serial, one stream: the drop-in calls back to back; isolates per-stage device times (latency view) ----
    serial_trace = []

    def serial_device(idx):
        for i in idx:
            e = [B.ev() for _ in range(4)]
            e[0].record()
            rgb = depth.process(frames[i % RING], H)
            e[1].record()
            d = depth.predict_depth(rgb)
            e[2].record()
            sbs = make_sbs_core(rgb, d, depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE, out_layout="HWC")
            e[3].record()
            serial_trace.append(e)
        return sbs

    def serial_e2e(idx):
        for i in idx:
            rgb = depth.process(host_np[i % 4], H)            # H2D of the pinned BGRA frame happens here (depth.py:547)
            d = depth.predict_depth(rgb)
            out = depth.make_sbs(rgb, d, depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE)   # D2H float32 HWC + sync
        return out

    out = serial_device(range(B.warmup))
    assert tuple(out.shape) == (H, 2 * W, 3)
    serial_trace.clear()
    n_serial = min(args.steps, 100)
    ms_serial = B.region(serial_device, n_serial)
    torch.cuda.synchronize()
    st = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in serial_trace])
    serial_stage_ms = dict(zip(["process", "predict_depth", "warp"], np.median(st, 0).tolist()))
    serial_e2e(range(B.warmup))
    ms_serial_e2e = B.region(serial_e2e, n_serial)

    legs = pipe_legs(B, depth, frames, host_frames, H, W, 1, args.slots, args.steps, [("e2e", torch.float32), ("e2e_u8", torch.uint8), ("e2e_nv12", torch.uint8),
                                                                                    ("e2e_jpeg", torch.uint8), ("e2e_jpeg_desktop", torch.uint8)])
    dv, e2e, e2e8, e2en = legs["device"], legs["e2e"], legs["e2e_u8"], legs["e2e_nv12"]
    host_bw = host_copy_bandwidth(B)

    # ---- rooflines (denominators: MEASURED_PEAKS.json, else the profiling guide's fallback) ----
    traffic = load_traffic()
    # warp kernel: fp16 CHW rgb in (6 B/px) + fp16 depth in (2 B/px) + fp32 HWC Full-SBS out (24 B/px); timed alone (serial leg: one
    # frame at a time on one stream, CUDA events on that stream) so the duration is the kernel's, not a share of a busy GPU
    warp_bytes = H * W * (6 + 2 + 24)
    warp_gbs = warp_bytes / (serial_stage_ms["warp"] * 1e-3) / 1e9
    Hm, Wm = 294, 518
    gflop = model_flops(cfg, Hm, Wm) / 1e9
    net_tflops_iso = gflop / (serial_stage_ms["predict_depth"] * 1e-3) / 1e3
    net_tflops = gflop * args.steps / (dv["ms"] * 1e-3) / 1e3
    roofline_warp = {"kernel": "warp_sbs kernel (1080p)", "bound": "hbm", "achieved": warp_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": warp_gbs / peaks["hbm_gbs"], "traffic": (traffic.get("warp_1080p") or {}).get("bytes"), "traffic_detail": traffic.get("warp_1080p"), "peak_source": src, "bytes_per_launch": warp_bytes,
                     "duration_ms": serial_stage_ms["warp"], "timed": "alone on the GPU (serial leg), CUDA events on its stream, median of %d" % n_serial,
                     "io": "rgb fp16 CHW + depth fp16 -> fp32 HWC Full-SBS"}
    roofline_net = {"kernel": "whole frame graph (resize + ViT-B + DPT on gemm_tc_kernel/tcgen05 + postprocess + warp), all kernels",
                    "bound": "tensor", "achieved": net_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": net_tflops / peaks["bf16_tflops_sustained"], "traffic": None, "peak_source": src, "gflop_per_frame": gflop,
                    "duration_ms": dv["ms"] / args.steps, "timed": "timed region / frames with %d frames in flight" % args.slots,
                    "isolated": {"achieved": net_tflops_iso, "frac": net_tflops_iso / peaks["bf16_tflops_sustained"],
                                 "duration_ms": serial_stage_ms["predict_depth"], "note": "one frame alone on the GPU: batch-1 is latency-bound"}}
    gemms = gemm_rooflines(dev, peaks, src)
    # the dominant kernel = the tcgen05 GEMM (one kernel template; ~60 % of a frame's GPU time): its four encoder-layer launches at
    # THIS workload's shapes, timed live in a graph of back-to-back launches on their own stream; achieved = algorithmic FLOPs per
    # launch / average launch duration
    own = [r for r in gemms if r["M"] == 778]
    fl = sum(2.0 * r["M"] * r["N"] * r["K"] for r in own)
    us = sum(r["duration_us"] for r in own)
    inf = gemm_in_flight(dev, peaks, src, streams=args.slots)
    # `roofline` = the kernel as the timed region runs it (`slots` frames in flight, throughput-policy tiles, sustained peak);
    # `isolated` = one launch alone on the GPU against the burst peak (the latency view)
    roofline = {"kernel": "gemm_tc_kernel (tcgen05), the 4 GEMMs of one ViT-B encoder layer at M = 778 (qkv, proj, fc1+GELU, fc2)", "bound": "tensor",
                "achieved": inf["achieved"], "peak": inf["peak"], "unit": "TFLOP/s", "frac": inf["frac"],
                "traffic": (traffic.get("gemm_m778") or {}).get("bytes"), "traffic_detail": traffic.get("gemm_m778"),
                "peak_source": inf["peak_source"], "flops_per_launch": inf["flops_per_launch"], "duration_us_per_launch": inf["duration_us_per_launch"],
                "launches": inf["launches"], "streams": inf["streams"], "timed": inf["timed"],
                "isolated": {"achieved": fl / us / 1e6, "peak": peaks["bf16_tflops"], "frac": fl / us / 1e6 / peaks["bf16_tflops"],
                             "duration_us_per_launch": us / len(own), "peak_source": src + " (burst: kernel timed alone)",
                             "timed": "live: graph of 20 back-to-back launches per shape on one stream, L2-warm",
                             "note": "batch-1 shapes fill 42-126 of 148 SMs for ~10 us: latency-bound"},
                "note": "duration_us_per_launch = wall time x streams / launches (launches of different frames overlap); the same kernel at batch 8 is in large4k.roofline"}

    line = {
        "metric": METRIC, "value": dv["fps"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": B.warmup,
        "ms_per_step": dv["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 operands / fp32 accumulate", "data": "synthetic",
        "repeats": dv["repeats"], "region_ms_min": dv["min"], "region_ms_max": dv["max"],
        "config": {"workload": WORKLOADS["base1080"][4] + ", per GPU",
                   "l2": f"ring of {RING} distinct frames ({RING * H * W * 4 / 1e6:.0f} MB) > 126 MB L2",
                   "parallelism": f"frames sharded x{world}; {args.slots} frames in flight per GPU; one process per GPU bound to its NUMA node",
                   "weights": "seeded random init, one NCCL broadcast at init" if world > 1 else "seeded random init",
                   "timing": "median of `repeats` regions of exactly `steps` frames (>= 1 s of device time in total)"},
        "e2e": {"value": e2e["fps"], "unit": "frames/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                "ms_per_step": e2e["ms"] / args.steps, "repeats": e2e["repeats"], "pcie_gbs_per_gpu": e2e["pcie_gbs_per_gpu"],
                "api": f"StereoPipeline({args.slots} frames in flight) over d2s_pipe_*: pinned BGRA frame -> process -> predict_depth -> make_sbs -> float32 HWC host frame",
                "limiter": "the device->host copy of the reference-faithful float32 frame (49.8 MB/frame); measured pinned-copy rates in host_copy"},
        "gpu_launches": dv["launches"], "clocks": legs["clocks"], "numa": B.numa, "host_copy": host_bw,
        "legs": {
            "e2e_u8": {"value": e2e8["fps"], "unit": "frames/s", "h2d_bytes_per_step": e2e8["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e8["d2h_bytes_per_step"],
                       "pcie_gbs_per_gpu": e2e8["pcie_gbs_per_gpu"],
                       "note": "same loop, uint8 HWC frame packed by the warp kernel (4x fewer bytes over PCIe); what streamer.set_frame encodes (streamer.py:250-256)"},
            "e2e_nv12": {"value": e2en["fps"], "unit": "frames/s", "h2d_bytes_per_step": e2en["h2d_bytes_per_step"], "d2h_bytes_per_step": e2en["d2h_bytes_per_step"],
                         "pcie_gbs_per_gpu": e2en["pcie_gbs_per_gpu"],
                         "note": "same loop, NV12 frame (libjpeg colour conversion + 4:2:0 on the device, 1.5 B/px): 8x fewer bytes than float32"},
            "e2e_jpeg": jpeg_leg(legs["e2e_jpeg"], note="same loop, the whole JPEG encode on the device (replaces cv2.imencode, streamer.py:250-256)"),
            "e2e_jpeg_desktop": jpeg_leg(legs["e2e_jpeg_desktop"]),
            "serial": {"note": "the drop-in calls one frame at a time on one stream (latency view); stage_ms = medians", "steps": n_serial,
                       "fps_device": world * n_serial / (ms_serial / 1e3), "fps_e2e": world * n_serial / (ms_serial_e2e / 1e3),
                       "ms_per_frame_device": ms_serial / n_serial, "ms_per_frame_e2e": ms_serial_e2e / n_serial, "stage_ms": serial_stage_ms},
            "stage_ms_in_flight": dv["stage_ms"],
            "kernels_n2_n3": extra_kernel_legs(B),
        },
        "roofline": roofline, "roofline_warp": roofline_warp, "roofline_net": roofline_net, "roofline_gemm": gemms,
    }
    del frames, host_frames, host_np
    depth.model_wraper = None
    engine.close()
    torch.cuda.empty_cache()

    def finalize(aborted=None):
        """compact copies inside the keys the driver's parser keeps (e2e / roofline / cpu_baseline), and a summary as the LAST key so
        the tail of the line carries the 1080p and the 4K numbers side by side.  Also what the watchdog prints if a later block stalls."""
        if aborted:
            line["aborted"] = aborted
        l4 = line.get("large4k")
        line["e2e"]["legs"] = {"e2e_u8": e2e8["fps"], "e2e_nv12": e2en["fps"], "serial_ms_per_frame_device": ms_serial / n_serial, "serial_ms_per_frame_e2e": ms_serial_e2e / n_serial}
        line["roofline"]["others"] = {"gemm_m778_alone_tensor_frac_of_burst": roofline["isolated"]["frac"], "warp_1080p_hbm_frac": roofline_warp["frac"], "frame_graph_tensor_frac_in_flight": roofline_net["frac"],
                                      "frame_graph_tensor_frac_alone": roofline_net["isolated"]["frac"]}
        if l4:
            line["e2e"]["legs"].update({"large4k_value": l4["value"], "large4k_e2e_fp32": l4["e2e"]["value"], "large4k_e2e_u8": l4["e2e_u8"]["value"], "large4k_e2e_nv12": l4["e2e_nv12"]["value"]})
            line["roofline"]["others"].update({"warp_4k_hbm_frac": l4["roofline_warp"]["frac"], "gemm_m6224_tensor_frac": l4["roofline"]["frac"],
                                               "large4k_step_graph_tensor_frac": l4["roofline_net"]["frac"]})
            if "config5" in l4 and "value" in l4["config5"]:
                line["e2e"]["legs"].update({"config5_value": l4["config5"]["value"], "config5_e2e_fp32": l4["config5"]["e2e"], "config5_e2e_u8": l4["config5"]["e2e_u8"]})
        if "cpu_baseline" in line and "cpu_baseline_1thread" in line:
            line["cpu_baseline"]["one_thread"] = line["cpu_baseline_1thread"]["value"]
            if l4 and "cpu_baseline" in l4:
                line["cpu_baseline"]["large4k_all_cores"] = l4["cpu_baseline"]["value"]
            if "config1" in line:
                line["cpu_baseline"]["config1_one_thread"] = line["config1"]["cpu_1thread"]["value"]
            if "reference_cuda" in line:
                line["cpu_baseline"]["reference_torch_cuda_same_gpu"] = line["reference_cuda"]["value"]
        if "vda1080" in line:
            line["e2e"]["legs"].update({"vda1080_value": line["vda1080"]["value"], "vda1080_e2e_fp32": line["vda1080"]["e2e"]})
        line["summary"] = {"base1080": {"value": line["value"], "e2e_fp32": line["e2e"]["value"], "e2e_u8": e2e8["fps"], "e2e_nv12": e2en["fps"],
                                        "e2e_jpeg_noise": legs["e2e_jpeg"]["fps"], "e2e_jpeg_desktop": legs["e2e_jpeg_desktop"]["fps"]},
                           "large4k": ({"value": l4["value"], "e2e_fp32": l4["e2e"]["value"], "e2e_u8": l4["e2e_u8"]["value"], "e2e_nv12": l4["e2e_nv12"]["value"],
                                        "e2e_jpeg_desktop": l4["e2e_jpeg_desktop"]["value"]} if l4 else None),
                           "vda1080": ({"value": line["vda1080"]["value"], "e2e_fp32": line["vda1080"]["e2e"]} if "vda1080" in line else None),
                           "reference_cuda_base1080": line.get("reference_cuda", {}).get("value"), "unit": "frames/s", "n_gpus": world}
        return line

    # the headline is measured: from here on a stalled block costs that block, not the line
    if WATCHDOG is not None:
        WATCHDOG.finalize = finalize
    # ================= the 4K half of the metric: configs[2] per GPU, configs[4] across GPUs =================
    if not args.no_large4k:
        if WATCHDOG is not None:
            WATCHDOG.enter("large4k (configs[2], configs[4])", True)
        line["large4k"] = large4k_block(B, args, publish=lambda d: line.__setitem__("large4k", d))
        if WATCHDOG is not None:
            WATCHDOG.enter("vda1080 (configs[3])", True)
        line["vda1080"] = vda1080_block(B, args)
    if WATCHDOG is not None:
        WATCHDOG.enter("reference arms / wrap-up", True)
    # ================= reference arms measured in the same run (rank 0, N = 1) =================
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        rfps, rdt = cuda_reference_fps(30, 5, dev)
        line["reference_cuda"] = {"value": rfps, "unit": "frames/s", "kind": "port", "workload": "base1080",
                                  "sample": f"30 frames after 5 warm-ups ({rdt:.2f} s), one frame at a time with H2D + .cpu() per frame",
                                  "what": "the reference's torch-CUDA path: HF module under fp16 autocast (cuBLAS/cuDNN/SDPA) + ATen resize/topk/conv2d/grid_sample (oracle/cuda_pipeline.py)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cfps, cdt = cpu_reference_fps(args.cpu_frames, 1, threads)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_frames} frames of the same 1080p workload after 1 warm-up ({cdt:.1f} s)"}
        c1, c1dt = cpu_reference_fps(2, 1, 1)
        line["cpu_baseline_1thread"] = {"value": c1, "unit": "frames/s", "cores": 1, "kind": "port",
                                        "sample": f"2 frames after 1 warm-up ({c1dt:.1f} s); torch.set_num_threads(1) is what the reference ships (depth.py:19)"}
        line["config1"] = config1_block(B, threads)
    finalize()
    if WATCHDOG is not None:        # the line goes out before the process group is torn down (a teardown that stalls cannot lose it)
        WATCHDOG.done = True
        if rank == 0:
            write_all(WATCHDOG.fd, json.dumps(line) + "\n")
    elif rank == 0:
        print(json.dumps(line))
    if world > 1:
        B.dist.destroy_process_group()
    return 0


def large4k_block(B, args, publish=None):
    """configs[2] per GPU: DA-V2-Large, 8 concurrent 4K streams, one frame of each per step, batched through the network (6224 token
    rows) -> 8 Full-SBS frames.  Under torchrun also configs[4]: 8 streams x 8 frames in total, stream s -> rank s mod G."""
    torch, dev, world, rank = B.torch, B.dev, B.world, B.rank
    from desktop2stereo_b200 import depth
    variant, h, w, S, desc = WORKLOADS["large4k"]
    engine, cfg = build_engine(variant, rank, world, dev)
    depth.init(engine=engine, device=dev)
    g = torch.Generator(device=dev).manual_seed(SEED + 100 + rank)
    RING = 2                                   # 2 x 8 x 33 MB = 531 MB of distinct frames > L2
    frames = [torch.randint(0, 256, (S, h, w, 4), generator=g, dtype=torch.uint8, device=dev) for _ in range(RING)]
    host_frames = [f.cpu().pin_memory() for f in frames]
    steps = max(4, min(args.steps // 8, 16))    # a step = one frame of each of the 8 streams
    slots = int(os.environ.get("D2S_BENCH_L4K_SLOTS", "2"))
    legs = pipe_legs(B, depth, frames, host_frames, h, w, S, slots, steps, [("e2e", torch.float32), ("e2e_u8", torch.uint8), ("e2e_nv12", torch.uint8),
                                                                            ("e2e_jpeg_desktop", torch.uint8)])
    dv, e2e, e2e8, e2en = legs["device"], legs["e2e"], legs["e2e_u8"], legs["e2e_nv12"]
    gflop = S * model_flops(cfg, 294, 518) / 1e9
    step_ms = dv["ms"] / steps
    net_tf = gflop / (step_ms * 1e-3) / 1e3
    gem = [r for r in gemm_rooflines(dev, B.peaks, B.peak_src, shapes="large") ]
    fl = sum(2.0 * r["M"] * r["N"] * r["K"] for r in gem); us = sum(r["duration_us"] for r in gem)
    traffic = load_traffic()
    out = {"workload": desc + ", per GPU", "value": dv["fps"], "unit": "frames/s", "steps": steps, "frames_per_step": S, "ms_per_step": step_ms,
           "repeats": dv["repeats"], "slots": slots, "gpu_launches": dv["launches"], "stage_ms_in_flight": dv["stage_ms"], "clocks": legs["clocks"],
           "l2": "2 batches of 8 distinct 4K frames (531 MB) > 126 MB L2",
           "e2e": {"value": e2e["fps"], "unit": "frames/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                   "pcie_gbs_per_gpu": e2e["pcie_gbs_per_gpu"], "api": "StereoPipeline(streams=8, 2 steps in flight): pinned 4K BGRA frames -> float32 HWC host frames",
                   "limiter": "PCIe: 33 MB in + 199 MB out per 4K frame"},
           "e2e_u8": {"value": e2e8["fps"], "unit": "frames/s", "h2d_bytes_per_step": e2e8["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e8["d2h_bytes_per_step"],
                      "pcie_gbs_per_gpu": e2e8["pcie_gbs_per_gpu"]},
           "e2e_nv12": {"value": e2en["fps"], "unit": "frames/s", "h2d_bytes_per_step": e2en["h2d_bytes_per_step"], "d2h_bytes_per_step": e2en["d2h_bytes_per_step"],
                        "pcie_gbs_per_gpu": e2en["pcie_gbs_per_gpu"]},
           "e2e_jpeg_desktop": jpeg_leg(legs["e2e_jpeg_desktop"]),
           "target": {"north_star": ">= 60 frames/s end-to-end 4K depth + Full-SBS on 1 x B200", "met_fp32": e2e["fps"] / world >= 60, "met_u8": e2e8["fps"] / world >= 60},
           "roofline": {"kernel": "gemm_tc_persistent_kernel (tcgen05, cta_group::2), the 4 GEMMs of one ViT-L encoder layer at M = 6224", "bound": "tensor",
                        "achieved": fl / us / 1e6, "peak": B.peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": fl / us / 1e6 / B.peaks["bf16_tflops"],
                        "traffic": (traffic.get("gemm_m6224") or {}).get("bytes"), "traffic_detail": traffic.get("gemm_m6224"), "peak_source": B.peak_src + " (burst)", "per_shape": gem,
                        "timed": "live: graph of 20 back-to-back launches per shape, CUDA events on their stream"},
           "roofline_net": {"kernel": "whole step graph (8 x resize + ViT-L + DPT batch 8 + 8 x postprocess + 8 x warp)", "bound": "tensor",
                            "achieved": net_tf, "peak": B.peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": net_tf / B.peaks["bf16_tflops_sustained"],
                            "gflop_per_step": gflop, "duration_ms": step_ms},
           "roofline_warp": warp_roofline_4k(B, traffic)}
    if publish is not None:      # the per-GPU block is complete: what follows (configs[4]) adds to it, and cannot take it away if it stalls
        publish(out)
    if world > 1:      # configs[4]: 8 streams in total, stream s -> rank s % G, 8 frames per stream
        if WATCHDOG is not None:
            WATCHDOG.enter("config5 (configs[4])", True)
        from desktop2stereo_b200.sharding import streams_for_rank
        mine = streams_for_rank(8, rank, world)
        sl = len(mine)
        f5 = [(f[:sl] if sl > 1 else f[0]).contiguous() for f in frames]      # one stream per GPU: plain [h, w, 4] frames
        h5 = [f.cpu().pin_memory() for f in f5]
        l5 = pipe_legs(B, depth, f5, h5, h, w, sl, 2, 8, [("e2e", torch.float32), ("e2e_u8", torch.uint8)])
        # every rank runs 8 steps of its own streams; whole job = 64 frames per region
        scale = 64.0 / (world * 8 * sl)
        out["config5"] = {"workload": "BASELINE.json configs[4]: 8 concurrent 4K streams x 8 frames, DA-V2-Large -> Full-SBS; stream s -> rank s mod G",
                          "scaling": "strong", "streams_per_gpu": sl, "frames_per_region": 64,
                          "value": l5["device"]["fps"] * scale, "e2e": l5["e2e"]["fps"] * scale, "e2e_u8": l5["e2e_u8"]["fps"] * scale, "unit": "frames/s",
                          "repeats": l5["device"]["repeats"]}
    else:
        out["config5"] = {"note": "at 1 GPU configs[4] (8 streams on one GPU) is this block's own workload"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cfps, cdt = cpu_reference_fps(2, 1, threads, variant, h, w)
        out["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": f"2 4K frames of DA-V2-Large after 1 warm-up ({cdt:.1f} s), one frame at a time"}
    depth.model_wraper = None
    engine.close()
    torch.cuda.empty_cache()
    return out


def vda1080_block(B, args, videos=4, encoder="vits"):
    """configs[3]: streaming Video-Depth-Anything, 1080p, 32-frame temporal window.  Frames of one video are sequential (the K/V rings
    of the four temporal modules), so the unit of parallelism is the VIDEO: `videos` independent videos per GPU, each on its own
    one-slot StereoPipeline (own CUDA stream, own window); across GPUs "replicas only" (video v -> rank v mod G, no collective).
    A step = one frame of every video."""
    import numpy as np
    torch, dev, world = B.torch, B.dev, B.world
    from desktop2stereo_b200 import depth
    from desktop2stereo_b200.engine import B200Engine
    from desktop2stereo_b200.pipeline import StereoPipeline
    from desktop2stereo_b200.synth import make_vda_state_dict
    h, w = 1080, 1920
    engine = B200Engine.from_vda_state_dict(make_vda_state_dict(encoder, SEED), encoder, dev, out_dtype=torch.float16)
    depth.init(engine=engine, device=dev)
    g = torch.Generator(device=dev).manual_seed(SEED + 200 + B.rank)
    RING = 24
    frames = [torch.randint(0, 256, (h, w, 4), generator=g, dtype=torch.uint8, device=dev) for _ in range(RING)]
    host_frames = [f.cpu().pin_memory() for f in frames[:4]]
    out = {"workload": WORKLOADS["vda1080"][4] + f" ({encoder}, model input 294x518), {videos} videos per GPU", "unit": "frames/s", "videos_per_gpu": videos}
    steps = max(8, min(args.steps, 40))
    for name, host in (("value", False), ("e2e", True)):
        pipes = [StereoPipeline(depth_slots=1, display_mode=DISPLAY_MODE, depth_ratio=DEPTH_RATIO) for _ in range(videos)]
        src = host_frames if host else frames

        def run(idx):
            for i in idx:
                for v, p in enumerate(pipes):
                    if p.pending:
                        p.result(host=host)
                    f = src[(i * videos + v) % len(src)]
                    p.submit_pinned(f) if host else p.submit_device(f)
            for p in pipes:
                while p.pending:
                    p.result(host=host)
        run(range(34))                      # past the 32-frame window: steady state
        t = B.timed(run, steps)
        out[name] = world * videos * steps / (t["ms"] / 1e3)
        out[name + "_repeats"] = t["repeats"]
        if not host:                        # one video alone: latency per frame
            one = pipes[0]

            def single(idx):
                for i in idx:
                    one.submit_device(frames[i % RING]); one.result(host=False)
            ms = B.region(single, 20)
            out["single_video_ms_per_frame"] = ms / 20
        for p in pipes:
            p.close()
    out["steps"] = steps
    depth.model_wraper = None
    engine.close()
    torch.cuda.empty_cache()
    return out


def warp_roofline_4k(B, traffic):
    """the warp kernel alone at 4K, IN FLIGHT over a ring of frames larger than L2 (so its writes leave the L2): CUDA events around a
    graph of 16 launches on 16 distinct input/output sets"""
    torch, dev = B.torch, B.dev
    from desktop2stereo_b200.stereo import make_sbs_core
    h, w, n = 2160, 3840, 6
    g = torch.Generator(device=dev).manual_seed(7)
    rgbs = [torch.randint(0, 256, (3, h, w), generator=g, device=dev, dtype=torch.uint8).half() for _ in range(n)]
    deps = [torch.rand((h, w), generator=g, device=dev).half() for _ in range(n)]
    outs = [torch.empty((h, 2 * w, 3), device=dev, dtype=torch.float32) for _ in range(n)]
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        for i in range(n):
            make_sbs_core(rgbs[i], deps[i], depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE, out_layout="HWC", out=outs[i])
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(n):
                make_sbs_core(rgbs[i], deps[i], depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE, out_layout="HWC", out=outs[i])
        gr.replay(); st.synchronize()
        s_, e_ = B.ev(), B.ev()
        s_.record(st); gr.replay(); gr.replay(); e_.record(st); st.synchronize()
    ms = s_.elapsed_time(e_) / (2 * n)
    nbytes = h * w * (6 + 2 + 24)
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "warp_sbs kernel (4K)", "bound": "hbm", "achieved": gbs, "peak": B.peaks["hbm_gbs"],
            "unit": "GB/s", "frac": gbs / B.peaks["hbm_gbs"], "traffic": (traffic.get("warp_4k") or {}).get("bytes"), "traffic_detail": traffic.get("warp_4k"), "bytes_per_launch": nbytes, "duration_ms": ms,
            "timed": f"live: {2 * n} launches over {n} distinct 4K frame sets ({n * nbytes / 1e6:.0f} MB > L2) in a CUDA graph, events on its stream",
            "io": "rgb fp16 CHW + depth fp16 -> fp32 HWC Full-SBS", "peak_source": B.peak_src}


def config1_block(B, threads):
    """BASELINE configs[0]: DA-V2-Small, one 518x518 frame -> Half-SBS: the reference's CPU path (1 thread as shipped, and all
    cores) next to this repo's serial drop-in calls on the GPU."""
    torch, dev = B.torch, B.dev
    import numpy as np
    from desktop2stereo_b200 import depth
    from desktop2stereo_b200.engine import B200Engine
    c1, d1 = cpu_reference_fps(2, 1, 1, "Small", 518, 518, "Half-SBS")
    ca, da = cpu_reference_fps(4, 1, threads, "Small", 518, 518, "Half-SBS")
    eng = B200Engine.from_hf_model(build_hf_model("Small"), dev)
    depth.init(engine=eng, device=dev)
    rng = np.random.default_rng(0)
    frame = torch.from_numpy(rng.integers(0, 256, (518, 518, 4), dtype=np.uint8)).pin_memory().numpy()

    def run(idx):
        for _ in idx:
            rgb = depth.process(frame, 518)
            out = depth.make_sbs(rgb, depth.predict_depth(rgb), display_mode="Half-SBS")
        return out
    assert run(range(5)).shape == (518, 518, 3)
    ms = B.region(run, 50)
    depth.model_wraper = None
    eng.close()
    return {"workload": "BASELINE.json configs[0]: DA-V2-Small, 518x518 frame -> Half-SBS (model input 518x518, 1370 tokens)",
            "cpu_1thread": {"value": c1, "unit": "frames/s", "cores": 1, "kind": "port", "sample": f"2 frames ({d1:.1f} s)"},
            "cpu_all_cores": {"value": ca, "unit": "frames/s", "cores": threads, "kind": "port", "sample": f"4 frames ({da:.1f} s)"},
            "b200_serial_e2e": {"value": 50 / (ms / 1e3), "unit": "frames/s", "note": "process -> predict_depth -> make_sbs, host frame in, float32 host frame out, one at a time"}}


def extra_kernel_legs(B):
    """the widened rows' kernels alone, CUDA events around 10 launches at 1080p: the occlusion-aware DIBR renderer
    on an edge-rich depth map (u8 frame + fp32 depth -> fp32 Full-SBS, 3 + 4 + 24 B/px) and the NV12 packer (3 B/px in, 1.5 B/px out)"""
    torch, dev = B.torch, B.dev
    from desktop2stereo_b200.stereo import make_sbs_dibr, rgb_to_nv12
    h, w = 1080, 1920
    g = torch.Generator(device=dev).manual_seed(3)
    rgb = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8, device=dev)
    yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    dep = (0.2 + 0.6 * ((xx // 97 + yy // 53) % 2)).float()
    out = torch.empty((h, 2 * w, 3), device=dev)
    frame8 = torch.randint(0, 256, (h, 2 * w, 3), generator=g, dtype=torch.uint8, device=dev)
    nv = torch.empty((h * 3 // 2, 2 * w), dtype=torch.uint8, device=dev)
    res = {}
    st = torch.cuda.Stream(dev)
    for name, fn, nbytes in (("dibr_1080p", lambda: make_sbs_dibr(rgb, dep, depth_ratio=2.0, display_mode="Full-SBS", out=out, out_layout="HWC"), h * w * (3 + 4 + 24)),
                             ("nv12_1080p_full_sbs", lambda: rgb_to_nv12(frame8, out=nv), h * 2 * w * 4.5)):
        with torch.cuda.stream(st):          # a CUDA graph of 10 launches: the host's launch cost is not what is measured
            for _ in range(3):
                fn()
            st.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(10):
                    fn()
            gr.replay(); st.synchronize()
            s_, e_ = B.ev(), B.ev()
            s_.record(st); gr.replay(); e_.record(st); st.synchronize()
        ms = s_.elapsed_time(e_) / 10
        res[name] = {"duration_us": ms * 1e3, "bytes_per_launch": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "peak_gbs": B.peaks["hbm_gbs"],
                     "frac": nbytes / (ms * 1e-3) / 1e9 / B.peaks["hbm_gbs"], "bound": "hbm", "timed": "graph of 10 launches, CUDA events on its stream, L2-warm"}
    res["dibr_1080p"]["note"] = "two passes: every pixel's 5 bilinear depth fetches + colour fetch, then the inpaint sweeps of the queued edge pixels densely; gather / latency-bound (L2-resident taps), edge-rich test map"
    return res


def load_traffic():
    """per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernels from this round's ncu --set full
    captures: profiles/traffic.json, written by tools/ncu_summary.py from the committed captures (never typed in here)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def gemm_rooflines(dev, peaks, src, shapes="base"):
    """The tcgen05 GEMM kernel alone, timed live with CUDA events around a graph of back-to-back launches (so that host launch
    cost is not what is measured): the headline workload's own shapes (M = 778 token rows: latency-bound) and the same layers at
    batch 8 (M = 6224: persistent kernel).  Peak = the measured cuBLAS bf16 BURST figure (a kernel timed in isolation)."""
    import torch
    from desktop2stereo_b200 import _lib
    L = _lib.lib()
    out = []
    table = {"base": [("qkv, batch 1 (ViT-B)", 778, 2304, 768, False, 0), ("proj (+residual stream)", 778, 768, 768, True, 0),
                      ("fc1 + GELU", 778, 3072, 768, False, 1), ("fc2 (+residual stream)", 778, 768, 3072, True, 0)],
             "large": [("qkv, batch 8 (ViT-L)", 6224, 3072, 1024, False, 0), ("proj (+residual stream)", 6224, 1024, 1024, True, 0),
                       ("fc1 + GELU", 6224, 4096, 1024, False, 1), ("fc2 (+residual stream)", 6224, 1024, 4096, True, 0)]}
    for (name, M, N, K, x32, act) in table[shapes]:
        A = torch.randn(M, K, device=dev).half(); B = torch.randn(N, K, device=dev).half() * (K ** -0.5); bias = torch.randn(N, device=dev)
        C = torch.empty(M, N, device=dev, dtype=torch.float16); X = torch.zeros(M, N, device=dev)
        st = torch.cuda.Stream(dev)
        iters = 20

        def call():
            _lib.check(L.d2s_debug_gemm(A.data_ptr(), B.data_ptr(), bias.data_ptr(), None if x32 else C.data_ptr(), M, N, K, act,
                                        X.data_ptr() if x32 else None, torch.cuda.current_stream(dev).cuda_stream))
        with torch.cuda.stream(st):
            for _ in range(3):
                call()
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(iters):
                    call()
            g.replay(); st.synchronize()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(st); g.replay(); e_.record(st); st.synchronize()
        us = s_.elapsed_time(e_) * 1e3 / iters
        tf = 2.0 * M * N * K / us / 1e6
        out.append({"kernel": "gemm_tc_kernel / gemm_tc_persistent_kernel (tcgen05)", "layer": name, "M": M, "N": N, "K": K, "bound": "tensor",
                    "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"], "duration_us": us,
                    "peak_source": src + " (burst)", "timed": "graph of %d back-to-back launches, CUDA events on its stream, L2-warm" % iters})
    return out


def gemm_in_flight(dev, peaks, src, shapes="base", streams=8, reps=5):
    """The dominant kernel as it runs in the headline timed region: `streams` frames in flight, each issuing the encoder layer's four
    GEMMs back to back with the throughput-policy tiles.  Every stream replays a CUDA graph of `reps` layers; wall time between two
    events that bracket all streams.  achieved = total algorithmic FLOPs / wall time = FLOPs per launch / (wall time x streams / launches)."""
    import torch
    from desktop2stereo_b200 import _lib
    L = _lib.lib()
    table = {"base": [(778, 2304, 768, False, 0), (778, 768, 768, True, 0), (778, 3072, 768, False, 1), (778, 768, 3072, True, 0)]}[shapes]
    _lib.check(L.d2s_debug_set_gemm_policy(1), "d2s_debug_set_gemm_policy")
    try:
        sts = [torch.cuda.Stream(dev) for _ in range(streams)]
        graphs, keep = [], []
        for st in sts:
            ops = []
            for (M, N, K, x32, act) in table:
                A = torch.randn(M, K, device=dev).half(); Bw = torch.randn(N, K, device=dev).half() * (K ** -0.5); bias = torch.randn(N, device=dev)
                C = torch.empty(M, N, device=dev, dtype=torch.float16); X = torch.zeros(M, N, device=dev)
                keep.append((A, Bw, bias, C, X))
                ops.append((A, Bw, bias, None if x32 else C, M, N, K, act, X if x32 else None))

            def layer(ops=ops):
                for (A, Bw, bias, C, M, N, K, act, X) in ops:
                    _lib.check(L.d2s_debug_gemm(A.data_ptr(), Bw.data_ptr(), bias.data_ptr(), C.data_ptr() if C is not None else None, M, N, K, act,
                                                X.data_ptr() if X is not None else None, torch.cuda.current_stream(dev).cuda_stream))
            with torch.cuda.stream(st):
                layer(); st.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    for _ in range(reps):
                        layer()
                g.replay(); st.synchronize()
            graphs.append(g)
        main = torch.cuda.current_stream(dev)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        s_.record(main)
        for st, g in zip(sts, graphs):
            st.wait_event(s_)
            with torch.cuda.stream(st):
                g.replay(); g.replay()
        for st in sts:
            main.wait_stream(st)
        e_.record(main)
        torch.cuda.synchronize(dev)
    finally:
        _lib.check(L.d2s_debug_set_gemm_policy(0), "d2s_debug_set_gemm_policy")
    us = s_.elapsed_time(e_) * 1e3
    launches = streams * 2 * reps * len(table)
    fl = streams * 2 * reps * sum(2.0 * M * N * K for (M, N, K, _, _) in table)
    tf = fl / us / 1e6
    return {"achieved": tf, "frac": tf / peaks["bf16_tflops_sustained"], "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "launches": launches,
            "streams": streams, "wall_us": us, "flops_per_launch": fl / launches, "duration_us_per_launch": us * streams / launches,
            "peak_source": src + " (sustained: a long run of overlapping launches)",
            "timed": f"live: {streams} streams x graphs of {2 * reps} encoder layers (4 GEMMs each, throughput-policy tiles), CUDA events around all streams"}


def load_peaks():
    peaks, src = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))); src = "measured"
    except Exception:
        pass
    return peaks, src


def build_engine(variant, rank, world, dev):
    """rank 0 packs the seeded weights, one NCCL broadcast over NVLink, every rank builds its own engine"""
    import torch
    from transformers import DepthAnythingConfig
    from desktop2stereo_b200 import sharding
    from desktop2stereo_b200.engine import B200Engine
    from desktop2stereo_b200.weights import config_from_hf, pack_state_dict
    blob = cfg_json = None
    if rank == 0:
        model = build_hf_model(variant)
        blob = pack_state_dict(model.state_dict(), config_from_hf(model.config))
        cfg_json = model.config.to_json_string()
        del model
    blob, cfg_json = sharding.broadcast_weights(blob, cfg_json, src=0, device=dev)
    cfg = config_from_hf(DepthAnythingConfig.from_dict(json.loads(cfg_json)))
    return B200Engine(blob, cfg, dev, out_dtype=torch.float16), cfg


def run_vda1080(args):
    """configs[3]: streaming Video-Depth-Anything at 1080p.  The temporal state makes consecutive frames of ONE video sequential,
    so the unit of parallelism is the video: `--slots` videos run on `--slots` CUDA streams of one GPU (one per GPU across GPUs:
    "replicas only").  One step = one frame of every video; value = frames/s over all videos.  Extra measurement."""
    import numpy as np
    import torch
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.engine import B200Engine
    from desktop2stereo_b200.prepost import PostProcessor, preprocess, process
    from desktop2stereo_b200.stereo import make_sbs_core
    from desktop2stereo_b200.synth import make_vda_state_dict
    h, w, S = 1080, 1920, args.slots
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    L = _lib.lib()
    engine = B200Engine.from_vda_state_dict(make_vda_state_dict(args.vda_encoder, SEED), args.vda_encoder, dev, out_dtype=torch.float16)
    g = torch.Generator(device=dev).manual_seed(SEED)
    RING = 24
    frames = [torch.randint(0, 256, (h, w, 4), generator=g, dtype=torch.uint8, device=dev) for _ in range(RING)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    posts = [PostProcessor() for _ in range(S)]
    outs = [torch.empty((h, 2 * w, 3), dtype=torch.float32, device=dev) for _ in range(S)]
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(i):
        for s_, st in enumerate(streams):
            with torch.cuda.stream(st):
                rgb = process(frames[(i * S + s_) % RING], h)
                x = preprocess(rgb, 518, 14)
                raw = engine(x)
                d = posts[s_](raw.reshape(raw.shape[-2:]), out_size=(h, w))
                make_sbs_core(rgb, d, depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE, out_layout="HWC", out=outs[s_])

    warmup = max(args.warmup, 3)
    for i in range(warmup + 33):            # past the 32-frame window: steady state
        step(i)
    torch.cuda.synchronize()
    # one video alone (latency view)
    s0, e0 = ev(), ev()
    with torch.cuda.stream(streams[0]):
        s0.record()
        for i in range(20):
            rgb = process(frames[i % RING], h); x = preprocess(rgb, 518, 14); raw = engine(x)
            d = posts[0](raw.reshape(raw.shape[-2:]), out_size=(h, w))
            make_sbs_core(rgb, d, depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE, out_layout="HWC", out=outs[0])
        e0.record()
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index or 0); clocks.start()
    l0 = L.d2s_launch_count()
    s, e = ev(), ev()
    s.record()
    for st in streams:
        st.wait_event(s)
    for i in range(args.steps):
        step(i)
    for st in streams:
        e.wait(st) if False else torch.cuda.current_stream(dev).wait_stream(st)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    line = {"metric": "end-to-end frames/sec (depth infer + SBS warp)", "value": S * args.steps / (ms / 1e3), "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands / fp32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOADS["vda1080"][4] + f" ({args.vda_encoder}, model input 294x518)", "parallelism": f"{S} videos on {S} CUDA streams",
                       "l2": f"ring of {RING} distinct frames (199 MB) > 126 MB L2"},
            "gpu_launches": int(L.d2s_launch_count() - l0), "clocks": clocks.stop(),
            "single_video": {"ms_per_frame": s0.elapsed_time(e0) / 20, "fps": 20 / (s0.elapsed_time(e0) / 1e3)},
            "state_bytes_per_video": None}
    print(json.dumps(line))
    return 0


def _main_with_clean_stdout():
    """Libraries (NCCL's version banner, torchrun's notices) write to fd 1; the contract is ONE JSON line on stdout.  Everything
    written while the benchmark runs is diverted to stderr, and only the result line goes to the real stdout."""
    import io
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    global WATCHDOG
    WATCHDOG = Watchdog()
    WATCHDOG.fd = real
    WATCHDOG.rank = int(os.environ.get("RANK", "0"))
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    rc = 1
    try:
        rc = main()
    finally:
        sys.stdout = old
        sys.stdout.flush()
        os.dup2(real, 1)
        os.close(real)
        for l in buf.getvalue().splitlines():
            if l.strip():
                (sys.stdout if l.lstrip().startswith("{") else sys.stderr).write(l + "\n")
        sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(_main_with_clean_stdout())
