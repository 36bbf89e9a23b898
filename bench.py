#!/usr/bin/env python
"""bench.py — end-to-end frames/s of the depth-inference + SBS-warp hot path (BASELINE.json metric).

Workload at every N: BASELINE.json configs[1] per GPU — Depth Anything V2-Base, 1080p BGRA frames, batch 1,
Full-SBS (model input 294x518, 778 tokens) — synthetic frames, seeded random-init weights (no network).  Frames shard
across ranks with no per-frame collective ("scaling": "weak"); the only collective is one NCCL broadcast of the packed
weight blob at init.

    python bench.py --gpus 1 --steps 200 --warmup 10          # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port), host cores
One JSON line on stdout (rank 0).  `value` = frames/s with frames resident in HBM; `e2e` = the same through the
reference-facing calls process -> predict_depth -> make_sbs with HOST buffers (H2D of the frame and D2H of the float32
SBS frame inside the timed region).
"""
from __future__ import annotations

import argparse
import json
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
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_hf_model(variant=VARIANT, seed=SEED):
    from desktop2stereo_b200.synth import make_hf_model   # seeded random-init weights (no checkpoints offline)
    return make_hf_model(variant, seed)


def cpu_reference_fps(n_frames, warmup, threads):
    """The reference's CPU path (oracle port: HF model under bf16 autocast + torch pre/post/warp) on the host cores."""
    import numpy as np
    import torch
    from oracle.cpu_pipeline import ReferenceCPUPipeline
    torch.set_num_threads(threads)
    pipe = ReferenceCPUPipeline(build_hf_model(), 518)
    rng = np.random.default_rng(SEED)
    frames = [rng.integers(0, 256, (H, W, 4), dtype=np.uint8) for _ in range(2)]
    for i in range(warmup):
        pipe.frame(frames[i % 2], DISPLAY_MODE, DEPTH_RATIO)
    t0 = time.perf_counter()
    for i in range(n_frames):
        out = pipe.frame(frames[i % 2], DISPLAY_MODE, DEPTH_RATIO)
    dt = time.perf_counter() - t0
    assert out.shape == (H, 2 * W, 3)
    return n_frames / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    fps, dt = cpu_reference_fps(args.steps, max(args.warmup, 1), threads)
    line = {
        "impl": "reference", "metric": "end-to-end frames/sec (depth infer + SBS warp)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 autocast (CPU)", "data": "synthetic",
        "config": {"workload": f"DA-V2-{VARIANT}, 1080p BGRA batch=1 -> {DISPLAY_MODE} (model input 294x518)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} frames of the same 1080p workload, 1 frame per step"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


WORKLOADS = {
    # name: (variant, frame h, frame w, frames per engine call, description)
    "base1080": ("Base", 1080, 1920, 1, "BASELINE.json configs[1]: DA-V2-Base, 1080p BGRA batch=1 -> Full-SBS (model input 294x518, 778 tokens)"),
    "large4k": ("Large", 2160, 3840, 8, "BASELINE.json configs[2]: DA-V2-Large, 4K BGRA batch=8 -> Full-SBS (model input 8 x 294x518, 6224 token rows)"),
    "vda1080": ("vits", 1080, 1920, 1, "BASELINE.json configs[3]: streaming Video-Depth-Anything, 1080p, 32-frame temporal window, one video per CUDA stream"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=3)
    ap.add_argument("--slots", type=int, default=8, help="frames in flight (CUDA streams) in the pipelined legs")
    ap.add_argument("--vda-encoder", default="vits", choices=["vits", "vitb", "vitl"])
    ap.add_argument("--workload", default="base1080", choices=sorted(WORKLOADS),
                    help="base1080 is the headline (configs[1]); large4k (configs[2]) is an extra measurement, not the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "large4k":
        return run_large4k(args)
    if args.workload == "vda1080":
        return run_vda1080(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from desktop2stereo_b200 import _lib, depth
    from desktop2stereo_b200.stereo import make_sbs_core
    L = _lib.lib()
    warmup = max(args.warmup, 3)
    engine, cfg = build_engine(VARIANT, rank, world, dev)
    depth.init(engine=engine, device=dev)

    # ---- synthetic frames: a ring larger than L2 so no timed iteration re-reads a cached frame ----
    RING = 24
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    frames = [torch.randint(0, 256, (H, W, 4), generator=g, dtype=torch.uint8, device=dev) for _ in range(RING)]
    host_frames = [f.cpu().pin_memory() for f in frames[:4]]   # pinned host copies for the end-to-end leg
    host_np = [t.numpy() for t in host_frames]

    from desktop2stereo_b200.pipeline import StereoPipeline
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        """`steps` frames through fn (a callable that consumes an iterable of frame indices), device-timed, max over ranks."""
        barrier(); torch.cuda.synchronize()
        s, e = ev(), ev()
        s.record()
        fn(range(steps))
        torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize(); barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def stage_stats(trace):
        """per-stage device time from the events recorded on each frame's stream: median (robust to a host hiccup) and mean"""
        torch.cuda.synchronize()
        st = np.array([list(e) if not hasattr(e[0], "elapsed_time") else [e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in trace])
        names = ["process", "predict_depth", "warp"]
        return dict(zip(names, np.median(st, 0).tolist())), dict(zip(names, st.mean(0).tolist()))

    # ---- (1) serial, one stream: the drop-in calls back to back; isolates per-stage device times ----
    serial_trace = []

    def serial_device(idx):
        for i in idx:
            e = [ev() for _ in range(4)]
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

    out = serial_device(range(warmup))
    assert tuple(out.shape) == (H, 2 * W, 3)
    serial_trace.clear()
    n_serial = min(args.steps, 100)
    ms_serial = timed(serial_device, n_serial)
    serial_stage_ms, serial_stage_mean = stage_stats(serial_trace)
    serial_e2e(range(warmup))
    ms_serial_e2e = timed(serial_e2e, n_serial)

    # ---- (2) pipelined: `slots` frames in flight on `slots` CUDA streams (what main.py's 3-thread loop does) ----
    pipe = StereoPipeline(depth_slots=args.slots, display_mode=DISPLAY_MODE, depth_ratio=DEPTH_RATIO)

    def pipe_device(idx):
        for _ in pipe.run((frames[i % RING] for i in idx), host=False):
            pass

    def pipe_e2e(idx, p=pipe):
        for res in p.run((host_frames[i % 4] for i in idx), host=True):   # pinned host frames -> H2D inside the timed region
            pass
        return res

    pipe_device(range(max(warmup, 2 * args.slots)))
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = L.d2s_launch_count()
    pipe.trace = []
    torch.cuda.profiler.start()          # no-op unless run under `ncu --profile-from-start off`
    ms = timed(pipe_device, args.steps)
    torch.cuda.profiler.stop()
    launches = L.d2s_launch_count() - launches0
    clk = clocks.stop()
    stage_ms, _ = stage_stats(pipe.trace)   # events recorded on each frame's stream INSIDE the timed region (streams overlap)
    pipe.trace = None
    fps = world * args.steps / (ms / 1e3)

    res = pipe_e2e(range(max(warmup, 2 * args.slots)))
    assert res.shape == (H, 2 * W, 3) and res.dtype == np.float32
    ms_e2e = timed(pipe_e2e, args.steps)
    fps_e2e = world * args.steps / (ms_e2e / 1e3)
    h2d, d2h = H * W * 4, H * 2 * W * 3 * 4

    # the same end-to-end loop with the 4x smaller u8 frame packed on the device (not the reference's return type)
    pipe8 = StereoPipeline(depth_slots=args.slots, display_mode=DISPLAY_MODE, depth_ratio=DEPTH_RATIO, out_dtype=torch.uint8)
    res8 = pipe_e2e(range(max(warmup, 2 * args.slots)), pipe8)
    assert res8.shape == (H, 2 * W, 3) and res8.dtype == np.uint8
    ms_e2e8 = timed(lambda idx: pipe_e2e(idx, pipe8), args.steps)

    # ---- rooflines (denominators: MEASURED_PEAKS.json, else the profiling guide's fallback) ----
    peaks, src = load_peaks()
    # warp kernel: fp16 CHW rgb in (6 B/px) + fp16 depth in (2 B/px) + fp32 HWC Full-SBS out (24 B/px); timed alone (serial leg:
    # one frame at a time on one stream, CUDA events on that stream) so the duration is the kernel's, not a share of a busy GPU
    warp_bytes = H * W * (6 + 2 + 24)
    warp_gbs = warp_bytes / (serial_stage_ms["warp"] * 1e-3) / 1e9
    # the network is replayed as ONE CUDA-graph launch per frame (~140 kernels, the tcgen05 GEMM is ~60 % of its time):
    #   isolated: duration of one launch alone on the GPU (serial leg) -> batch-1 latency view
    #   in the timed region: `slots` launches overlap, so GPU time per launch = timed region / launches
    Hm, Wm = 294, 518
    gflop = model_flops(cfg, Hm, Wm) / 1e9
    net_tflops_iso = gflop / (serial_stage_ms["predict_depth"] * 1e-3) / 1e3
    net_tflops = gflop * args.steps / (ms * 1e-3) / 1e3
    roofline_warp = {"kernel": "warp_sbs_fast_kernel", "bound": "hbm", "achieved": warp_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": warp_gbs / peaks["hbm_gbs"], "traffic": 16.63e6 + 1.22e6, "traffic_note": "dram__bytes_read + dram__bytes_write of one isolated launch under ncu --set full (profiles/r1_warp_v4_ncu_full.txt): the 49.8 MB written stay in the 126 MB L2 during an isolated replay",
                     "peak_source": src, "bytes_per_launch": warp_bytes,
                     "duration_ms": serial_stage_ms["warp"], "timed": "alone on the GPU (serial leg), CUDA events on its stream, median of %d" % n_serial,
                     "io": "rgb fp16 CHW + depth fp16 -> fp32 HWC Full-SBS"}
    roofline_net = {"kernel": "depth network, one graph launch per frame (preprocess + ViT-B + DPT on gemm_tc_kernel/tcgen05 + postprocess)",
                    "bound": "tensor", "achieved": net_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": net_tflops / peaks["bf16_tflops_sustained"], "traffic": None, "peak_source": src, "gflop_per_launch": gflop,
                    "duration_ms": ms / args.steps, "timed": "timed region / launches with %d launches in flight" % args.slots,
                    "isolated": {"achieved": net_tflops_iso, "frac": net_tflops_iso / peaks["bf16_tflops_sustained"],
                                 "duration_ms": serial_stage_ms["predict_depth"], "note": "one launch alone on the GPU: batch-1 is latency-bound"}}

    line = {
        "metric": "end-to-end frames/sec (depth infer + SBS warp)", "value": fps, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 operands / fp32 accumulate", "data": "synthetic",
        "config": {"workload": f"DA-V2-{VARIANT}, 1080p BGRA batch=1 -> {DISPLAY_MODE} (model input 294x518, 778 tokens), per GPU",
                   "l2": f"ring of {RING} distinct frames ({RING * H * W * 4 / 1e6:.0f} MB) > 126 MB L2", "parallelism": f"frames sharded x{world}; {args.slots} frames in flight per GPU",
                   "weights": "seeded random init, one NCCL broadcast at init" if world > 1 else "seeded random init"},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "pcie_gbs": (h2d + d2h) * fps_e2e / world / 1e9,
                "api": f"StereoPipeline({args.slots} frames in flight): pinned BGRA frame -> process -> predict_depth -> make_sbs -> float32 HWC host frame",
                "note": "bounded by the device->host copy of the reference-faithful float32 frame (49.8 MB/frame)"},
        "e2e_u8": {"value": world * args.steps / (ms_e2e8 / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": H * 2 * W * 3,
                   "note": "same loop, uint8 HWC frame packed by the warp kernel (4x fewer bytes over PCIe); not the reference's return dtype"},
        "gpu_launches": int(launches), "clocks": clk, "stage_ms": stage_ms,
        "serial": {"note": "same calls, one frame at a time on one stream (latency view); stage_ms = medians", "steps": n_serial,
                   "fps_device": world * n_serial / (ms_serial / 1e3), "fps_e2e": world * n_serial / (ms_serial_e2e / 1e3),
                   "ms_per_frame_device": ms_serial / n_serial, "ms_per_frame_e2e": ms_serial_e2e / n_serial, "stage_ms": serial_stage_ms,
                   "stage_ms_mean": serial_stage_mean},
        "roofline": roofline_net, "roofline_warp": roofline_warp, "roofline_net": roofline_net,
        "roofline_gemm": gemm_rooflines(dev, peaks, src),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cfps, cdt = cpu_reference_fps(args.cpu_frames, 1, threads)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_frames} frames of the same 1080p workload after 1 warm-up ({cdt:.1f} s)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def gemm_rooflines(dev, peaks, src):
    """The tcgen05 GEMM kernel alone, timed live with CUDA events around a graph of back-to-back launches (so that host launch
    cost is not what is measured): the headline workload's own shapes (M = 778 token rows: latency-bound) and the same layers at
    batch 8 (M = 6224: persistent kernel).  Peak = the measured cuBLAS bf16 BURST figure (a kernel timed in isolation)."""
    import torch
    from desktop2stereo_b200 import _lib
    L = _lib.lib()
    out = []
    for (name, M, N, K, x32) in [("qkv, batch 1", 778, 2304, 768, False), ("fc2 (+residual stream), batch 1", 778, 768, 3072, True),
                                 ("qkv, batch 8 (ViT-L)", 6224, 3072, 1024, False), ("fc1, batch 8 (ViT-L)", 6224, 4096, 1024, False)]:
        A = torch.randn(M, K, device=dev).half(); B = torch.randn(N, K, device=dev).half() * (K ** -0.5); bias = torch.randn(N, device=dev)
        C = torch.empty(M, N, device=dev, dtype=torch.float16); X = torch.zeros(M, N, device=dev)
        st = torch.cuda.Stream(dev)
        iters = 20

        def call():
            _lib.check(L.d2s_debug_gemm(A.data_ptr(), B.data_ptr(), bias.data_ptr(), None if x32 else C.data_ptr(), M, N, K, 0,
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


def run_large4k(args):
    """configs[2]: DA-V2-Large, 8 x 4K frames per engine call -> 8 Full-SBS frames.  Extra measurement (M = 6224 token rows is
    where the tcgen05 GEMM is tensor-bound rather than latency-bound); one step = one batch of 8 frames."""
    import numpy as np
    import torch
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.prepost import PostProcessor, preprocess, process
    from desktop2stereo_b200.stereo import make_sbs_core
    variant, h, w, B, desc = WORKLOADS["large4k"]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    engine, cfg = build_engine(variant, 0, 1, dev)
    L = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(SEED)
    RING = 2                                   # 2 x 8 x 33 MB = 531 MB of distinct frames > L2
    frames = [[torch.randint(0, 256, (h, w, 4), generator=g, dtype=torch.uint8, device=dev) for _ in range(B)] for _ in range(RING)]
    posts = [PostProcessor() for _ in range(B)]    # 8 concurrent streams: one EMA state each
    batch = torch.empty((B, 3, 294, 518), dtype=torch.float32, device=dev)
    outs = [torch.empty((h, 2 * w, 3), dtype=torch.float32, device=dev) for _ in range(B)]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    trace = []

    def step(i):
        fr = frames[i % RING]
        e = [ev() for _ in range(4)]
        e[0].record()
        rgbs = [process(f, h) for f in fr]
        for b, rgb in enumerate(rgbs):
            preprocess(rgb, 518, 14, out=batch[b:b + 1])
        e[1].record()
        raw = engine(batch)
        e[2].record()
        for b in range(B):
            d = posts[b](raw[b], out_size=(h, w))
            make_sbs_core(rgbs[b], d, depth_ratio=DEPTH_RATIO, display_mode=DISPLAY_MODE, out_layout="HWC", out=outs[b])
        e[3].record()
        trace.append(e)

    warmup = max(args.warmup, 3)
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    trace.clear()
    clocks = ClockSampler(dev.index or 0); clocks.start()
    l0 = L.d2s_launch_count()
    s, e = ev(), ev()
    s.record()
    for i in range(args.steps):
        step(i)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    clk = clocks.stop()
    st = np.median(np.array([[t[j].elapsed_time(t[j + 1]) for j in range(3)] for t in trace]), 0)
    peaks, src = load_peaks()
    gflop = B * model_flops(cfg, 294, 518) / 1e9
    net_tf = gflop / (st[1] * 1e-3) / 1e3
    line = {"metric": "end-to-end frames/sec (depth infer + SBS warp)", "value": B * args.steps / (ms / 1e3), "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands / fp32 accumulate", "data": "synthetic",
            "config": {"workload": desc, "l2": "2 batches of 8 distinct 4K frames (531 MB) > 126 MB L2", "parallelism": "one stream, batch of 8 per engine call"},
            "gpu_launches": int(L.d2s_launch_count() - l0), "clocks": clk,
            "stage_ms": {"process+preprocess x8": float(st[0]), "engine (batch 8)": float(st[1]), "postprocess+warp x8": float(st[2])},
            "roofline": {"kernel": "depth network, one graph launch per batch of 8 (ViT-L + DPT on gemm_tc_kernel/tcgen05)", "bound": "tensor",
                         "achieved": net_tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": net_tf / peaks["bf16_tflops_sustained"],
                         "traffic": None, "peak_source": src, "gflop_per_launch": gflop, "duration_ms": float(st[1])},
            "roofline_warp": {"kernel": "warp_sbs_fast_kernel x8 (+ postprocess x8)", "bound": "hbm", "bytes_per_launch": h * w * 32,
                              "note": "stage time covers 8 post-process chains and 8 warp launches"}}
    print(json.dumps(line))
    return 0


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
