"""GPU parity of the warp kernel (through the C ABI) against the oracle — bit-exact colours and indices —
and against ATen's own CUDA kernels driven by the torch restatement of make_sbs_core."""
import os

import numpy as np
import pytest
import torch

from oracle import warp as owarp

pytestmark = pytest.mark.gpu
MODES = ["Full-SBS", "Half-SBS", "Full-TAB", "Half-TAB"]
DTS = ["float32", "float16", "bfloat16"]
TDT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def _run(dev, rgb_u8, depth, dt, mode, fill, conv, ratio, gather, rgb_dtype=None, **kw):
    from desktop2stereo_b200.stereo import make_sbs_core
    rgb_t = torch.from_numpy(rgb_u8).to(dev)
    if rgb_dtype is not None:
        rgb_t = rgb_t.to(rgb_dtype)
    dep_t = torch.from_numpy(depth).to(dev).to(TDT[dt])
    return make_sbs_core(rgb_t, dep_t, 0.064, ratio, mode, fill, conv, gather=gather, **kw)


def test_golden_cases_bit_exact(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "warp.npz"))
    for i in range(int(g["n_cases"])):
        k = f"c{i:03d}"
        h, w, mode, fill, dt, gather = (int(v) for v in g[k + "_meta"])
        ipd, ratio, conv = (float(v) for v in g[k + "_par"])
        for rgb_dtype in (None, TDT[DTS[dt]]):  # u8 source and the reference's float source
            out = _run(cuda_device, g[k + "_rgb"], g[k + "_depth"], DTS[dt], MODES[mode], bool(fill), conv, ratio, bool(gather),
                       rgb_dtype=rgb_dtype)
            ref = g[k + "_out"]
            assert tuple(out.shape) == ref.shape
            assert np.array_equal(out.float().cpu().numpy(), ref), (k, MODES[mode], fill, DTS[dt], gather)


@pytest.mark.parametrize("h,w", [(270, 480), (135, 241), (518, 518)])
def test_indices_and_colours_vs_oracle(cuda_device, h, w):
    rng = np.random.default_rng(h * 7 + w)
    for mode in MODES:
        for gather in (False, True):
            for dt in DTS:
                rgb = rng.integers(0, 256, (3, h, w)).astype(np.uint8)
                dep = rng.random((h, w)).astype(np.float32)
                dep = torch.from_numpy(dep).to(TDT[dt]).float().numpy()
                out, il, ir = _run(cuda_device, rgb, dep, dt, mode, True, 0.5, 4.0, gather, return_indices=True)
                o, ol, orr = owarp.make_sbs_core_oracle(rgb.astype(np.float32), dep, 0.064, 4.0, mode, True, 0.5,
                                                        depth_dtype=dt, gather=gather, return_indices=True)
                assert np.array_equal(il.cpu().numpy(), ol), (mode, gather, dt, "left indices")
                assert np.array_equal(ir.cpu().numpy(), orr), (mode, gather, dt, "right indices")
                assert np.array_equal(out.float().cpu().numpy(), o), (mode, gather, dt)


def test_vs_aten_cuda_kernels(cuda_device):
    """Same frames through ATen's CUDA linspace/grid_sample/gather/adaptive_avg_pool (what the reference runs on CUDA).
    Tolerance: 1e-3 * 255 on colours (BASELINE north star); in practice the two agree to the last bit."""
    rng = np.random.default_rng(11)
    worst = 0.0
    for (h, w) in [(1080, 1920), (360, 641)]:
        for mode in MODES:
            for dt in ("float32", "float16"):
                rgb = rng.integers(0, 256, (3, h, w)).astype(np.uint8)
                dep = torch.from_numpy(rng.random((h, w)).astype(np.float32)).to(TDT[dt])
                out = _run(cuda_device, rgb, dep.float().numpy(), dt, mode, False, 0.0, 2.0, False)
                ref = owarp.make_sbs_core_torch(torch.from_numpy(rgb).to(cuda_device).to(TDT[dt]), dep.to(cuda_device),
                                                0.064, 2.0, mode, False, 0.0)
                err = (out - ref.float()).abs().max().item()
                worst = max(worst, err)
                assert err <= 1e-3 * 255, (h, w, mode, dt, err)
    print("max |kernel - ATen CUDA| =", worst)


def test_layouts_and_dtypes(cuda_device):
    """HWC / CHW / BGRA sources and f32 / f16 / u8 outputs all describe the same picture."""
    from desktop2stereo_b200.stereo import make_sbs_core
    rng = np.random.default_rng(5)
    h, w = 121, 200
    rgb = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    dep = torch.from_numpy(rng.random((h, w)).astype(np.float32)).to(cuda_device)
    chw = torch.from_numpy(rgb).to(cuda_device).permute(2, 0, 1).contiguous()
    base = make_sbs_core(chw, dep, display_mode="Full-SBS")
    hwc = make_sbs_core(torch.from_numpy(rgb).to(cuda_device), dep, display_mode="Full-SBS", rgb_layout="HWC", out_layout="HWC")
    assert torch.equal(hwc.permute(2, 0, 1), base)
    bgra = np.concatenate([rgb[..., ::-1], np.full((h, w, 1), 255, np.uint8)], axis=2)
    b = make_sbs_core(torch.from_numpy(bgra).to(cuda_device), dep, display_mode="Full-SBS", rgb_layout="BGRA")
    assert torch.equal(b, base)
    u8 = make_sbs_core(chw, dep, display_mode="Full-SBS", out_dtype=torch.uint8, out_layout="HWC")
    assert torch.equal(u8.permute(2, 0, 1), base.round().to(torch.uint8))
    f16 = make_sbs_core(chw, dep, display_mode="Half-SBS", out_dtype=torch.float16)
    assert torch.equal(f16, make_sbs_core(chw, dep, display_mode="Half-SBS").half())


def test_lowres_depth_fused_upsample(cuda_device):
    """Passing the model-resolution depth map == upsampling it first (depth.py:1998-2004) then warping."""
    from desktop2stereo_b200.stereo import make_sbs_core
    import torch.nn.functional as F
    rng = np.random.default_rng(9)
    h, w = 1080, 1920
    rgb = torch.from_numpy(rng.integers(0, 256, (3, h, w)).astype(np.uint8)).to(cuda_device)
    for dt in (torch.float32, torch.float16):
        low = torch.from_numpy(rng.random((294, 518)).astype(np.float32)).to(cuda_device).to(dt)
        full = F.interpolate(low[None, None], size=(h, w), mode="bilinear", align_corners=False)[0, 0]
        a = make_sbs_core(rgb, low, display_mode="Full-SBS")
        b = make_sbs_core(rgb, full, display_mode="Full-SBS")
        # the upsample is floating point: one fp16 ulp of depth moves a pixel by < 0.02 px
        assert (a - b).abs().max().item() <= (1e-3 * 255 if dt == torch.float32 else 3.0), dt
        assert (a - b).abs().mean().item() < 0.05


def test_properties_full_size(cuda_device):
    """Size-independent properties at 4K: zero shift at depth == convergence is the identity picture in the gather
    branch; Half-SBS equals the pair-mean of Full-SBS; left/right swap under shift negation."""
    from desktop2stereo_b200.stereo import make_sbs_core
    g = torch.Generator(device="cpu").manual_seed(0)
    h, w = 2160, 3840
    rgb = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8).to(cuda_device)
    dep = torch.rand((h, w), generator=g).to(cuda_device)
    flat = torch.full((h, w), 0.5, device=cuda_device)
    ident = make_sbs_core(rgb, flat, convergence=0.5, display_mode="Full-SBS", gather=True, out_dtype=torch.float32)
    assert torch.equal(ident[:, :, :w], rgb.float()) and torch.equal(ident[:, :, w:], rgb.float())
    full = make_sbs_core(rgb, dep, display_mode="Full-SBS")
    half = make_sbs_core(rgb, dep, display_mode="Half-SBS")
    assert torch.equal(half, (full[:, :, 0::2] + full[:, :, 1::2]) * 0.5)
    tab = make_sbs_core(rgb, dep, display_mode="Full-TAB")
    assert torch.equal(tab[:, :h], full[:, :, :w]) and torch.equal(tab[:, h:], full[:, :, w:])
    # negating the disparity swaps the eyes: depth' = 2*conv - depth
    sw = make_sbs_core(rgb, dep, convergence=0.25, display_mode="Full-SBS", gather=True, out_dtype=torch.float32)
    sw2 = make_sbs_core(rgb, 0.5 - dep, convergence=0.25, display_mode="Full-SBS", gather=True, out_dtype=torch.float32)
    assert (sw[:, :, :w] != sw2[:, :, w:]).float().mean().item() < 1e-3


def test_make_sbs_host_signature(cuda_device):
    """make_sbs keeps depth.py:2186's contract: ndarray HWC or tensor CHW in, float32 HWC ndarray out."""
    from desktop2stereo_b200.stereo import make_sbs
    rng = np.random.default_rng(2)
    h, w = 90, 160
    rgb = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    dep = rng.random((h, w)).astype(np.float32)
    a = make_sbs(rgb, torch.from_numpy(dep).to(cuda_device), display_mode="Half-SBS")
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == (h, w, 3)
    o = owarp.make_sbs_core_oracle(rgb.transpose(2, 0, 1).astype(np.float32), dep, display_mode="Half-SBS")
    assert np.array_equal(a, o.transpose(1, 2, 0))
    b = make_sbs(torch.from_numpy(rgb).permute(2, 0, 1).to(cuda_device).half(), torch.from_numpy(dep).to(cuda_device).half(),
                 display_mode="Full-SBS", fill_16_9=True)
    assert b.shape == (h, 2 * w, 3)


def test_fast_path_equals_generic_and_oracle(cuda_device):
    """The smem-staged source-centric kernel and the generic output-centric kernel are the same function, bit for bit —
    including depth outside [0,1], whose taps leave the staged window and fall back to global loads."""
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.stereo import make_sbs_core
    force_generic = _lib.lib().d2s_debug_force_generic_warp
    rng = np.random.default_rng(21)
    for (h, w) in [(1080, 1920), (33, 70), (64, 1030), (37, 1296)]:
        rgb = rng.integers(0, 256, (3, h, w)).astype(np.uint8)
        for scale, conv, ratio in [(1.0, 0.0, 2.0), (1.0, 0.5, 4.0), (6.0, 0.0, 4.0)]:   # scale 6: depth in [-2.5, 3.5]
            dep = ((rng.random((h, w)).astype(np.float32) - 0.4) * scale).astype(np.float32)
            for mode in ("Full-SBS", "Full-TAB", "Half-SBS"):
                for dt in (torch.float16, torch.float32):
                    r = torch.from_numpy(rgb).to(cuda_device)
                    r = r.half() if dt == torch.float16 else r
                    d = torch.from_numpy(dep).to(cuda_device).to(dt)
                    for layout, odt in (("CHW", torch.float32), ("HWC", torch.float32), ("HWC", torch.uint8), ("CHW", torch.float16)):
                        force_generic(0)
                        fast = make_sbs_core(r, d, 0.064, ratio, mode, False, conv, out_layout=layout, out_dtype=odt)
                        force_generic(1)
                        try:
                            gen = make_sbs_core(r, d, 0.064, ratio, mode, False, conv, out_layout=layout, out_dtype=odt)
                        finally:
                            force_generic(0)
                        assert torch.equal(fast, gen), (h, w, scale, mode, dt, layout, odt)
                    fast = make_sbs_core(r, d, 0.064, ratio, mode, False, conv)
            if h < 100:
                o = owarp.make_sbs_core_oracle(rgb.astype(np.float32), d.float().cpu().numpy(), 0.064, ratio, "Half-SBS", False, conv,
                                               depth_dtype="float32")
                assert np.array_equal(fast.cpu().numpy(), o)
