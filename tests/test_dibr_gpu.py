"""d2s_make_sbs_dibr (the reference viewer's occlusion-aware DIBR shader, viewer.py:386-631, as a CUDA kernel) through the C ABI
against oracle/dibr_oracle.c: BIT-EXACT (strict fp32, same operation order) in every display mode, source dtype and layout;
the feathering path (powf) within 1e-5.  Plus size-independent properties at 4K."""
import numpy as np
import pytest
import torch

from oracle import dibr

_scene = dibr.synthetic_scene

pytestmark = pytest.mark.gpu
MODES = ["Full-SBS", "Half-SBS", "Full-TAB", "Half-TAB"]


def _run(dev, rgb, depth, mode, **kw):
    from desktop2stereo_b200.stereo import make_sbs_dibr
    out = make_sbs_dibr(torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev), display_mode=mode, rgb_layout="HWC", out_layout="HWC", **kw)
    return out.cpu().numpy()


@pytest.mark.parametrize("mode", MODES)
def test_dibr_bit_exact_vs_oracle(cuda_device, mode):
    for (h, w, seed) in [(90, 160, 1), (67, 131, 2), (128, 96, 3)]:
        rgb, depth = _scene(seed, h, w)
        depth = (depth + 0.05 * np.random.default_rng(seed).random((h, w))).astype(np.float32)     # texture on both sides of the edges
        for kw in (dict(ipd_uv=0.064, depth_ratio=3.0, convergence=0.2), dict(ipd_uv=0.1, depth_ratio=6.0, convergence=0.0, roll=0.2),
                   dict(ipd_uv=0.064, depth_ratio=2.0, convergence=0.5, corner_radius=0.08, search_radius=20, depth_tolerance=0.02, blur_radius=1.5)):
            want = dibr.make_sbs_dibr_oracle(rgb, depth, mode, **kw)
            got = _run(cuda_device, rgb, depth, mode, **kw)
            assert got.shape == want.shape and got.dtype == np.float32
            assert np.array_equal(got, want), (mode, h, w, kw, float(np.abs(got - want).max()))
    # the scene really exercises the inpaint path
    _, _, cl, cr = dibr.eye_views(rgb, depth, display_mode=mode, return_conf=True, ipd_uv=0.064, depth_ratio=3.0)
    assert (cl > 0.001).mean() > 0.005 and (cr > 0.001).mean() > 0.005


def test_dibr_two_pass_equals_single_pass(cuda_device):
    """the dense second pass over the queued edge pixels produces the same frame as the one-pass kernel, bit for bit"""
    from desktop2stereo_b200.stereo import make_sbs_dibr
    for (h, w, seed) in [(360, 640, 7), (135, 241, 8)]:
        rgb, depth = _scene(seed, h, w)
        r, d = torch.from_numpy(rgb).to(cuda_device), torch.from_numpy(depth).to(cuda_device)
        for mode in MODES:
            a = make_sbs_dibr(r, d, depth_ratio=4.0, display_mode=mode, rgb_layout="HWC", two_pass=True)
            b = make_sbs_dibr(r, d, depth_ratio=4.0, display_mode=mode, rgb_layout="HWC", two_pass=False)
            assert torch.equal(a, b), (h, w, mode)


def test_dibr_feather_dtypes_layouts(cuda_device):
    from desktop2stereo_b200.stereo import make_sbs_dibr
    h, w = 72, 128
    rgb, depth = _scene(5, h, w)
    kw = dict(ipd_uv=0.064, depth_ratio=3.0, convergence=0.1)
    want = dibr.make_sbs_dibr_oracle(rgb, depth, "Full-SBS", feather_enabled=True, feather_width=0.1, **kw)
    got = _run(cuda_device, rgb, depth, "Full-SBS", feather_enabled=True, feather_width=0.1, **kw)
    assert np.abs(got - want).max() <= 1e-5 * 255 * 4          # powf: a few ulp between libm and CUDA
    base = torch.from_numpy(dibr.make_sbs_dibr_oracle(rgb, depth, "Half-SBS", **kw)).to(cuda_device)
    r_hwc = torch.from_numpy(rgb).to(cuda_device)
    d = torch.from_numpy(depth).to(cuda_device)
    chw = make_sbs_dibr(r_hwc.permute(2, 0, 1).contiguous(), d, display_mode="Half-SBS", **kw)                  # planar u8 in, planar f32 out
    assert torch.equal(chw.permute(1, 2, 0), base)
    f16src = make_sbs_dibr(r_hwc.permute(2, 0, 1).contiguous().half(), d, display_mode="Half-SBS", **kw)        # process()'s fp16 tensor
    assert torch.equal(f16src, chw)
    bgra = torch.cat([r_hwc.flip(-1), torch.full((h, w, 1), 255, dtype=torch.uint8, device=cuda_device)], -1)   # the captured frame itself
    assert torch.equal(make_sbs_dibr(bgra, d, display_mode="Half-SBS", rgb_layout="BGRA", **kw), chw)
    u8 = make_sbs_dibr(r_hwc, d, display_mode="Half-SBS", rgb_layout="HWC", out_layout="HWC", out_dtype=torch.uint8, **kw)
    assert torch.equal(u8, base.round().to(torch.uint8))
    # fp16 depth (what predict_depth returns) == the same values as fp32
    d16 = d.half()
    assert torch.equal(make_sbs_dibr(r_hwc, d16, display_mode="Half-SBS", rgb_layout="HWC", **kw),
                       make_sbs_dibr(r_hwc, d16.float(), display_mode="Half-SBS", rgb_layout="HWC", **kw))


def test_dibr_properties_4k(cuda_device):
    """size-independent properties at 3840x2160: zero eye separation = the frame itself; TAB = SBS eyes; a flat depth map has no
    disocclusions, so the frame is a pure horizontal resample (rows independent)"""
    from desktop2stereo_b200.stereo import make_sbs_dibr
    g = torch.Generator(device="cpu").manual_seed(0)
    h, w = 2160, 3840
    rgb = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8).to(cuda_device)
    dep = torch.rand((h, w), generator=g).to(cuda_device)
    ident = make_sbs_dibr(rgb, dep, ipd_uv=0.0, display_mode="Full-SBS")
    # (the shader's border alpha, smoothstep(-0.001, 0.001, uv), fades the outermost 0.001 uv = 3.84 px at 4K: compare the interior;
    #  texel centres are hit to ~1e-4 px, i.e. a fraction of a grey level on a noise image)
    b = 8
    assert (ident[:, b:-b, b:w - b] - rgb.float()[:, b:-b, b:-b]).abs().max().item() <= 0.5 and torch.equal(ident[:, :, :w], ident[:, :, w:])
    assert ident[:, 0, :w].max().item() < rgb.float()[:, 0].max().item()
    full = make_sbs_dibr(rgb, dep, depth_ratio=2.0, display_mode="Full-SBS")
    tab = make_sbs_dibr(rgb, dep, depth_ratio=2.0, display_mode="Full-TAB")
    assert torch.equal(tab[:, :h], full[:, :, :w]) and torch.equal(tab[:, h:], full[:, :, w:])
    assert torch.isfinite(full).all() and full.min().item() >= 0 and full.max().item() <= 255
