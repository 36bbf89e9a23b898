"""GPU-box micro-benchmark: the tcgen05 GEMM at the encoder-layer shapes (bench.py's gemm_rooflines, stand-alone)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda:0")
peaks, src = bench.load_peaks()
for shapes in ("base", "large"):
    rows = bench.gemm_rooflines(dev, peaks, src, shapes=shapes)
    fl = sum(2.0 * r["M"] * r["N"] * r["K"] for r in rows); us = sum(r["duration_us"] for r in rows)
    for r in rows:
        print(f'{r["layer"]:32s} M={r["M"]:5d} N={r["N"]:5d} K={r["K"]:5d}  {r["duration_us"]:8.2f} us  {r["achieved"]:7.1f} TF/s  {r["frac"]:.3f}', flush=True)
    print(f'  layer aggregate: {fl / us / 1e6:7.1f} TF/s  {fl / us / 1e6 / peaks["bf16_tflops"]:.3f}')
