"""Kernel-level numerics: the tcgen05 GEMM, the implicit-GEMM 3x3 conv and the fused attention, each against a plain
PyTorch fp32 reference of the same op on the same fp16-rounded operands (tolerances: fp16 output rounding)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ACT = {"none": 0, "gelu": 1, "relu": 2}


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _act(x, name):
    return {"none": lambda t: t, "gelu": F.gelu, "relu": F.relu}[name](x)


@pytest.mark.parametrize("M,N,K,act", [
    (128, 128, 64, "none"), (128, 128, 256, "none"), (778, 2304, 768, "none"), (778, 3072, 768, "gelu"),
    (300, 96, 768, "relu"), (1000, 32, 576, "none"), (64, 64, 640, "none"), (1554, 384, 384, "none"),
    (6224, 1024, 4096, "none"), (37, 48, 128, "none"),
    (6224, 1024, 512, "gelu"), (6224, 3072, 256, "relu"), (5000, 2048, 320, "none"),     # enough tiles for the 128 x 256 tile variant
])
def test_gemm_tcgen05(cuda_device, M, N, K, act):
    from desktop2stereo_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N * 7 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(cuda_device)
    Bw = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).half().to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    Cc = torch.full((M, N), float("nan"), dtype=torch.float16, device=cuda_device)
    _lib.check(_lib.lib().d2s_debug_gemm(A.data_ptr(), Bw.data_ptr(), bias.data_ptr(), Cc.data_ptr(), M, N, K, ACT[act], None,
                                         _stream(cuda_device)), "d2s_debug_gemm")
    ref = _act(A.float() @ Bw.float().t() + bias, act)
    err = (Cc.float() - ref).abs().max().item()
    assert torch.isfinite(Cc).all()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), (M, N, K, err)


def test_gemm_residual_stream_wide_tiles(cuda_device):
    """The 128 x 256 tile variant on the fp32-stream epilogue (M large enough for the heuristic to pick it)."""
    from desktop2stereo_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(6)
    M, N, K = 6224, 1024, 1024
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(cuda_device)
    Bw = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).half().to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    X = torch.randn(M, N, generator=g).to(cuda_device)
    ref = X + (A.float() @ Bw.float().t() + bias)
    _lib.check(_lib.lib().d2s_debug_gemm(A.data_ptr(), Bw.data_ptr(), bias.data_ptr(), None, M, N, K, 0, X.data_ptr(),
                                         _stream(cuda_device)), "d2s_debug_gemm")
    assert (X - ref).abs().max().item() <= 1e-3


def test_gemm_residual_stream(cuda_device):
    """x32 += A W^T + b (the proj / fc2 epilogue), twice, to check the read-modify-write."""
    from desktop2stereo_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 778, 768, 3072
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(cuda_device)
    Bw = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).half().to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    X = torch.randn(M, N, generator=g).to(cuda_device)
    ref = X + 2 * (A.float() @ Bw.float().t() + bias)
    for _ in range(2):
        _lib.check(_lib.lib().d2s_debug_gemm(A.data_ptr(), Bw.data_ptr(), bias.data_ptr(), None, M, N, K, 0, X.data_ptr(),
                                             _stream(cuda_device)), "d2s_debug_gemm")
    assert (X - ref).abs().max().item() <= 1e-3


@pytest.mark.parametrize("B,H,W,Cin,N,act", [
    (1, 21, 37, 64, 64, "none"), (2, 11, 19, 128, 128, "relu"), (1, 84, 148, 128, 128, "none"), (1, 42, 74, 192, 64, "none"),
    (1, 50, 70, 64, 32, "relu"), (1, 5, 7, 64, 64, "none"), (1, 16, 8, 64, 64, "none"),
])
def test_conv3x3_implicit_gemm(cuda_device, B, H, W, Cin, N, act):
    from desktop2stereo_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(B + H * 3 + W * 5 + Cin)
    x = (torch.randn(B, Cin, H, W, generator=g)).half().to(cuda_device)
    w = (torch.randn(N, Cin, 3, 3, generator=g) * (1.0 / (9 * Cin) ** 0.5)).half().to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res = torch.randn(B, H, W, N, generator=g).half().to(cuda_device)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_k = w.permute(0, 2, 3, 1).reshape(N, 9 * Cin).contiguous()     # k = (ky*3+kx)*C + c
    out = torch.full((B, H, W, N), float("nan"), dtype=torch.float16, device=cuda_device)
    out_relu = torch.empty_like(out)
    _lib.check(_lib.lib().d2s_debug_conv3x3(x_nhwc.data_ptr(), w_k.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, Cin, N, ACT[act],
                                            res.data_ptr(), out_relu.data_ptr(), _stream(cuda_device)), "d2s_debug_conv3x3")
    ref = _act(F.conv2d(x.float(), w.float(), bias, padding=1), act).permute(0, 2, 3, 1) + res.float()
    assert torch.isfinite(out).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-3 * max(1.0, ref.abs().max().item()), err
    assert torch.equal(out_relu, F.relu(out))


@pytest.mark.parametrize("impl", ["tcgen05", "mma"])
@pytest.mark.parametrize("B,N,heads", [(1, 778, 6), (2, 1370, 2), (1, 36, 2), (3, 64, 1), (1, 65, 12), (8, 778, 16), (1, 128, 1), (1, 129, 3)])
def test_attention(cuda_device, monkeypatch, B, N, heads, impl):
    monkeypatch.setenv("D2S_ATTN", impl)   # tcgen05 (attention_tc.cu, the default) | mma (the mma.sync kernel)
    from desktop2stereo_b200 import _lib
    D = heads * 64
    g = torch.Generator(device="cpu").manual_seed(N + heads)
    qkv = torch.randn(B, N, 3 * D, generator=g).half().to(cuda_device)
    out = torch.full((B, N, D), float("nan"), dtype=torch.float16, device=cuda_device)
    _lib.check(_lib.lib().d2s_debug_attention(qkv.data_ptr(), out.data_ptr(), B, N, D, heads, _stream(cuda_device)), "d2s_debug_attention")
    q, k, v = (t.float().view(B, N, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(B, N, D)
    assert torch.isfinite(out).all()
    assert (out.float() - ref).abs().max().item() <= 4e-3
