#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM through d2s_debug_gemm: back-to-back launches (L2-hot, constant smem carve-out)
vs. launches interleaved with a zero-smem kernel, timed with CUDA events.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from desktop2stereo_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream(dev).cuda_stream


def bench(M, N, K, x32=False, interleave=False, iters=50):
    A = torch.randn(M, K, device=dev).half()
    B = torch.randn(N, K, device=dev).half()
    bias = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev, dtype=torch.float16)
    X = torch.zeros(M, N, device=dev)
    y = torch.zeros(1024, device=dev)
    def call():
        _lib.check(L.d2s_debug_gemm(A.data_ptr(), B.data_ptr(), bias.data_ptr(), None if x32 else C.data_ptr(), M, N, K, 0,
                                    X.data_ptr() if x32 else None, st()))
        if interleave:
            y.add_(1.0)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()          # replay from a graph so that host launch cost is not what is measured
    with torch.cuda.graph(g):
        for _ in range(iters):
            call()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / iters


for (M, N, K, x32) in [(778, 2304, 768, False), (778, 3072, 768, False), (778, 768, 768, True), (778, 768, 3072, True),
                        (128, 128, 768, False), (128, 128, 64, False)]:
    a = bench(M, N, K, x32)
    b = bench(M, N, K, x32, interleave=True)
    fl = 2 * M * N * K
    print(f"M={M} N={N} K={K} x32={x32}: back-to-back {a:7.1f} us ({fl / a / 1e6:7.1f} TFLOP/s)   interleaved with a 0-smem kernel {b:7.1f} us")
# tile shape / scheduling at tensor-bound sizes
CASES = (("128x128", dict(BN256=0, PAIR=0, PERSIST=0)), ("128x128 persist", dict(BN256=0, PAIR=0, PERSIST=1)),
         ("128x256", dict(BN256=2, PAIR=0, PERSIST=0)), ("128x256 persist", dict(BN256=2, PAIR=0, PERSIST=1)),
         ("pair256", dict(BN256=2, PAIR=2, PERSIST=0, BN256_STAGES=3)), ("pair256 persist", dict(BN256=2, PAIR=2, PERSIST=1)),
         ("default", dict()))
for (M, N, K, x32) in [(6224, 3072, 1024, False), (6224, 4096, 1024, False), (6224, 1024, 4096, True), (6224, 1024, 1024, True), (8192, 8192, 8192, False)]:
    fl = 2 * M * N * K
    res = []
    for name, env in CASES:
        for k in ("BN256", "PAIR", "PERSIST", "BN256_STAGES"):
            os.environ.pop("D2S_GEMM_" + k, None)
        for k, v in env.items():
            os.environ["D2S_GEMM_" + k] = str(v)
        a = bench(M, N, K, x32, iters=20)
        res.append(f"{name} {a:6.1f} us {fl / a / 1e6:5.0f} TF/s")
    print(f"M={M} N={N} K={K} x32={x32}:\n    " + "\n    ".join(res))
