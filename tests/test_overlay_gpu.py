"""GPU parity of d2s_overlay_fps (through the C ABI and the host mirror desktop2stereo_b200/overlay.py) against the oracle
and the reference goldens: bit-exact in every dtype / layout, including the reference's every-10th-call text cache."""
import os

import numpy as np
import pytest
import torch

from oracle.gen_golden import OVERLAY_CASES, overlay_rgb
from oracle.overlay import OverlayOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", OVERLAY_CASES, ids=lambda c: f"ov{c[0]}_{c[1]}x{c[2]}_{c[3]}")
def test_overlay_matches_reference_golden(cuda_device, golden_dir, case):
    from desktop2stereo_b200 import overlay
    seed, H, W, dt, fps_seq = case
    gold = np.load(os.path.join(golden_dir, "overlay.npz"))[f"ov{seed}"]
    base = overlay_rgb(seed, H, W)
    rgb = torch.from_numpy(base).to(cuda_device, getattr(torch, dt))
    keep = rgb.clone()
    overlay.reset_cache()
    oracle = OverlayOracle()
    ch, cw = min(H, 64), min(W, 420)
    for i, fps in enumerate(fps_seq):
        out = overlay.overlay_fps(rgb, fps)
        assert out.dtype == rgb.dtype and out.data_ptr() != rgb.data_ptr()
        o = out.float().cpu().numpy()
        assert np.array_equal(o[:, :ch, :cw].astype(np.uint8), gold[i]), (seed, i)
        assert np.array_equal(o, oracle(base, fps))
    assert torch.equal(rgb, keep)              # the caller's frame is not modified (the reference returns a new tensor)
    overlay.reset_cache()


@pytest.mark.parametrize("layout,dt", [("HWC", torch.uint8), ("HWC", torch.float32), ("CHW", torch.uint8)])
def test_overlay_layouts_inplace(cuda_device, layout, dt):
    from desktop2stereo_b200 import overlay
    H, W = 135, 240
    base = overlay_rgb(9, H, W)
    t = torch.from_numpy(base).to(cuda_device, dt)
    if layout == "HWC":
        t = t.permute(1, 2, 0).contiguous()
    overlay.reset_cache()
    out = overlay.overlay_fps(t, 72.5, layout=layout, inplace=True)
    assert out.data_ptr() == t.data_ptr()
    got = out.float().cpu().numpy()
    if layout == "HWC":
        got = got.transpose(2, 0, 1)
    assert np.array_equal(got, OverlayOracle()(base, 72.5))
    overlay.reset_cache()


def test_make_sbs_with_fps(cuda_device):
    """make_sbs(fps=...) == make_sbs of the overlaid frame (depth.py:2217-2219)."""
    from desktop2stereo_b200 import overlay
    from desktop2stereo_b200.stereo import make_sbs
    H, W = 120, 200
    rgb = torch.from_numpy(overlay_rgb(3, H, W)).to(cuda_device, torch.float16)
    dep = torch.rand(H, W, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(0)).half()
    overlay.reset_cache()
    a = make_sbs(rgb, dep, display_mode="Full-SBS", fps=50.0).copy()
    overlay.reset_cache()
    b = make_sbs(overlay.overlay_fps(rgb, 50.0), dep, display_mode="Full-SBS").copy()
    c = make_sbs(rgb, dep, display_mode="Full-SBS").copy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    overlay.reset_cache()
