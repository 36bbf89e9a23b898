"""End-to-end parity of the drop-in interface (process -> predict_depth -> make_sbs) against a golden produced by the
unmodified reference on the same seeded frame and weights (tests/golden/e2e.npz, fp32, autocast off, EMA off)."""
import os

import numpy as np
import pytest
import torch

from oracle.gen_golden import E2E, TINY, synth_frame
from oracle.ref_harness import make_hf_model

pytestmark = pytest.mark.gpu


def test_drop_in_pipeline_vs_reference_golden(cuda_device, golden_dir):
    from desktop2stereo_b200 import depth
    g = np.load(os.path.join(golden_dir, "e2e.npz"))
    model = make_hf_model("Small", E2E["seed"], TINY)
    depth.init(model, device=cuda_device, depth_resolution=E2E["depth_resolution"], fp16=False)
    frame = synth_frame(E2E["seed"], E2E["h"], E2E["w"], 4)
    rgb = depth.process(frame, E2E["h"])
    assert rgb.dtype == torch.float32 and tuple(rgb.shape) == (3, E2E["h"], E2E["w"])
    d = depth.predict_depth(rgb, use_temporal_smooth=False)
    assert tuple(d.shape) == (E2E["h"], E2E["w"])
    ref_d = torch.from_numpy(g["depth"])
    # The post-processed depth lives in [0,1]; the 2/98-percentile stretch amplifies the network's fp16 rounding.  Bound = what
    # the reference's own CUDA numerics score on this frame: the same weights under fp16 autocast (depth.py:1763-1781) through
    # the restated pre/post-processing (oracle/prepost.py, torch ops = the reference's ops), against the same fp32 golden.
    from oracle import prepost as opp
    err = (d.float().cpu() - ref_d).abs()
    mg = model.to(cuda_device)
    with torch.no_grad():
        x = opp.normalise_input(opp.resize_patch_aligned(rgb[None], E2E["depth_resolution"], 14))
        with torch.autocast("cuda", dtype=torch.float16):
            raw16 = mg(pixel_values=x).predicted_depth
        ref16 = opp.upsample_depth(opp.post_process_depth(raw16[0], 0.05, 4.0), E2E["h"], E2E["w"])
    err16 = (ref16.float().cpu() - ref_d).abs()
    print("post-processed depth: engine max err", err.max().item(), "mean", err.mean().item(),
          "| reference fp16 path max", err16.max().item(), "mean", err16.mean().item())
    assert err.max().item() <= 1.25 * err16.max().item() and err.mean().item() <= 1.25 * err16.mean().item()
    for mode in ("Half-SBS", "Full-SBS"):
        sbs = depth.make_sbs(rgb, d, ipd_uv=0.064, depth_ratio=4.0, convergence=0.0, display_mode=mode)
        ref = g["sbs_" + mode].astype(np.float32)
        assert sbs.dtype == np.float32 and sbs.shape == ref.shape
        e = np.abs(sbs - ref)
        # a depth error of a few 1e-3 moves a pixel by < 0.01 px; on this smooth frame that is well under one grey level
        print(mode, "sbs max err", e.max(), "mean", e.mean())
        assert e.max() <= 1.0 and e.mean() <= 0.05
    # same depth into the reference's own warp == bit-exact picture (isolates the warp from the network's fp16 noise)
    from oracle import warp as owarp
    sbs = depth.make_sbs(rgb, ref_d.to(cuda_device), depth_ratio=4.0, display_mode="Full-SBS")
    o = owarp.make_sbs_core_oracle(rgb.cpu().numpy(), g["depth"], depth_ratio=4.0, display_mode="Full-SBS")
    assert np.array_equal(sbs, o.transpose(1, 2, 0))


def test_temporal_smoothing_state(cuda_device):
    """DepthStabilizer semantics (depth.py:1865-1887): first frame passes through, later frames are EMA'd per stream."""
    from desktop2stereo_b200 import depth
    model = make_hf_model("Small", 2, TINY)
    depth.init(model, device=cuda_device, depth_resolution=70)
    f0, f1 = synth_frame(1, 90, 160, 4), synth_frame(2, 90, 160, 4)
    a0 = depth.predict_depth(depth.process(f0, 90), use_temporal_smooth=False)
    a1 = depth.predict_depth(depth.process(f1, 90), use_temporal_smooth=False)
    depth.depth_stabilizer.reset()
    b0 = depth.predict_depth(depth.process(f0, 90))
    b1 = depth.predict_depth(depth.process(f1, 90))
    # first frame passes through the stabiliser unchanged: the engine is deterministic, so bit-identical
    assert torch.equal(a0, b0)
    mix = 0.9 * a0.float() + 0.1 * a1.float()
    assert (b1.float() - mix).abs().max().item() <= 2e-3     # EMA runs on the low-res fp16 map before the upsample (2 fp16 ulp at 1.0)
    assert not torch.equal(a1, b1)
