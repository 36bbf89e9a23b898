"""Device JPEG encoder: per-kernel times on the BASELINE output sizes (CUDA events, graph-free), stream sizes, cv2 host time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import jpeg as oj
from desktop2stereo_b200.stereo import JpegEncoder

dev = torch.device("cuda:0")
for (h, w, name) in [(1080, 3840, "1080p Full-SBS"), (2160, 7680, "4K Full-SBS")]:
    for content in ("desktop", "noise"):
        img = oj.desktop_like(h, w, 1) if content == "desktop" else np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
        t = torch.from_numpy(img).to(dev)
        for ri in [int(x) for x in os.environ.get("JPEG_RI", "1,2,4,8,16").split(",")]:
            enc = JpegEncoder(h, w, dev, quality=90, restart_interval=ri)
            for _ in range(3): enc.encode(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): enc.encode(t)
            e1.record(); torch.cuda.synchronize()
            n = int(enc.size.item())
            print(f"{name} {content} ri={ri}: {e0.elapsed_time(e1) / 20 * 1e3:.0f} us/frame, {n} bytes ({n / (h * w):.3f} B/px)", flush=True)
        t0 = time.perf_counter(); oj.encode_cv2(img, 90, 0); print(f"   cv2.imencode host: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
