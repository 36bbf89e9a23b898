#!/usr/bin/env python
"""One tiny tcgen05 GEMM launch (for ncu source-level inspection of the fixed per-launch cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from desktop2stereo_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 64)))
A = torch.randn(M, K, device=dev).half(); B = torch.randn(N, K, device=dev).half(); bias = torch.randn(N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.float16)
for _ in range(4):
    _lib.check(L.d2s_debug_gemm(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 0, None,
                                torch.cuda.current_stream(dev).cuda_stream))
torch.cuda.synchronize()
print("ok", (C.float() - (A.float() @ B.float().t() + bias)).abs().max().item())
