"""GPU-box micro-benchmark of the warp kernel alone: 1080p / 4K, fp32 / u8 outputs, L2-warm (one frame set) and in flight
(frame sets larger than L2), CUDA events around a CUDA graph of launches."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from desktop2stereo_b200.stereo import make_sbs_core

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.0


def run(h, w, n, odt, mode="Full-SBS", rgb_dt=torch.float16):
    g = torch.Generator(device=dev).manual_seed(7)
    rgbs = [torch.randint(0, 256, (3, h, w), generator=g, device=dev, dtype=torch.uint8).to(rgb_dt) for _ in range(n)]
    deps = [torch.rand((h, w), generator=g, device=dev).half() for _ in range(n)]
    ow = 2 * w if mode == "Full-SBS" else w
    outs = [torch.empty((h, ow, 3), device=dev, dtype=odt) for _ in range(n)]
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        for i in range(n):
            make_sbs_core(rgbs[i], deps[i], display_mode=mode, out_layout="HWC", out=outs[i])
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(n):
                make_sbs_core(rgbs[i], deps[i], display_mode=mode, out_layout="HWC", out=outs[i])
        gr.replay(); st.synchronize()
        reps = max(2, 24 // n)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(st)
        for _ in range(reps):
            gr.replay()
        e_.record(st); st.synchronize()
    us = s_.elapsed_time(e_) * 1e3 / (reps * n)
    nbytes = h * w * (rgbs[0].element_size() * 3 + 2) + h * ow * 3 * outs[0].element_size()
    print(f"{h}x{w} {mode} rgb {str(rgb_dt)[6:]} -> {str(odt)[6:]} sets={n}: {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s  {nbytes / us / 1e3 / peak:.3f} of {peak:.0f}", flush=True)


for (h, w) in [(1080, 1920), (2160, 3840)]:
    for n in (1, 8 if h == 1080 else 6):
        run(h, w, n, torch.float32)
        run(h, w, n, torch.uint8)
run(2160, 3840, 6, torch.float32, "Half-SBS")
run(2160, 3840, 6, torch.uint8, "Full-SBS", torch.uint8)
