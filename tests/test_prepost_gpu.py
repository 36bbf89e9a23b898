"""GPU parity of process / preprocess / postprocess (through the C ABI) against the oracle and the goldens.
Integer work (shapes, channel swizzle) is bit-exact; floating-point stages carry explicit tolerances."""
import os

import numpy as np
import pytest
import torch

from oracle import prepost as opp
from oracle.gen_golden import synth_depth, synth_frame

pytestmark = pytest.mark.gpu
DT = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}


def test_model_input_shape_bit_exact(cuda_device, golden_dir):
    from desktop2stereo_b200.prepost import model_input_shape
    pre = np.load(os.path.join(golden_dir, "pre.npz"))
    for h, w, target, nh, nw in pre["shapes"]:
        assert model_input_shape(int(h), int(w), int(target)) == (int(nh), int(nw))
    rng = np.random.default_rng(0)
    for _ in range(500):
        h, w, t = int(rng.integers(14, 4400)), int(rng.integers(14, 7700)), int(rng.choice([518, 336, 294, 224, 70]))
        assert model_input_shape(h, w, t) == opp.model_input_shape(h, w, t), (h, w, t)


def test_process_swizzle_bit_exact(cuda_device, golden_dir):
    from desktop2stereo_b200.prepost import process
    pre = np.load(os.path.join(golden_dir, "pre.npz"))
    _, h, w, ch, th = (int(v) for v in pre["proc0_meta"])
    frame = synth_frame(0, h, w, ch)
    for name, dt in (("float32", torch.float32), ("float16", torch.float16)):
        out = process(frame, th, dtype=dt)
        assert out.dtype == dt and tuple(out.shape) == (3, h, w)
        assert np.array_equal(out.float().cpu().numpy(), pre[f"proc0_{name}"].astype(np.float32))
    # ragged sizes / 3-channel / full HD
    for (hh, ww, c) in [(1080, 1920, 4), (7, 13, 3), (33, 5, 4)]:
        f = synth_frame(5, hh, ww, c)
        out = process(torch.from_numpy(f).to(cuda_device), hh)
        assert torch.equal(out.cpu(), opp.process_cuda_branch(torch.from_numpy(f), hh, torch.float16))


def test_process_downscale(cuda_device, golden_dir):
    """bilinear + antialias downscale (depth.py:560-566): fp32 golden within 1e-3 of 255 (kernel accumulates fp32)."""
    from desktop2stereo_b200.prepost import process
    pre = np.load(os.path.join(golden_dir, "pre.npz"))
    for seed in (1, 2):
        _, h, w, ch, th = (int(v) for v in pre[f"proc{seed}_meta"])
        out = process(synth_frame(seed, h, w, ch), th, dtype=torch.float32)
        ref = pre[f"proc{seed}_float32"]
        assert tuple(out.shape) == ref.shape
        assert np.abs(out.cpu().numpy() - ref).max() <= 2e-3, seed


def test_preprocess_vs_golden_and_oracle(cuda_device, golden_dir):
    from desktop2stereo_b200.prepost import preprocess
    pre = np.load(os.path.join(golden_dir, "pre.npz"))
    for seed in (10, 11, 12, 13, 14):
        _, h, w, target = (int(v) for v in pre[f"pre{seed}_meta"])
        frame = synth_frame(seed, h, w, 3)  # BGR
        ref = pre[f"pre{seed}_input"]
        # BGR HWC capture layout read in place
        x = preprocess(torch.from_numpy(frame).to(cuda_device), target, layout="BGR")
        assert tuple(x.shape) == (1,) + ref.shape
        err = np.abs(x.cpu().numpy()[0] - ref).max()
        assert err <= 2e-5, (seed, err)   # fp32 resample: a few ulp at |x| <= 2.7
        # CHW fp16 tensor (what process() returns)
        chw = torch.from_numpy(frame[..., ::-1].copy()).permute(2, 0, 1).contiguous().to(cuda_device).half()
        x2 = preprocess(chw, target)
        assert torch.equal(x2, x)
    # full-size: 4K BGRA -> 294x518, against the oracle run on the same GPU (ATen CUDA bicubic-AA)
    f4k = synth_frame(20, 2160, 3840, 4)
    x = preprocess(torch.from_numpy(f4k).to(cuda_device), 518, layout="BGRA")
    t = torch.from_numpy(f4k[..., 2::-1].copy()).permute(2, 0, 1).unsqueeze(0).to(cuda_device)
    ref = opp.normalise_input(opp.resize_patch_aligned(t, 518, 14))
    assert tuple(x.shape) == (1, 3, 294, 518)
    assert (x - ref).abs().max().item() <= 5e-5


def test_postprocess_vs_golden(cuda_device, golden_dir):
    from desktop2stereo_b200.prepost import PostProcessor
    post = np.load(os.path.join(golden_dir, "post.npz"))
    fg, aa = (float(v) for v in post["fg_aa"])
    # 1 ulp of the compute dtype for the point ops; powf differs from the CPU's by <= 2 ulp fp32
    tol = {0: 2e-6, 1: 2e-3, 2: 1.6e-2}
    for i in range(int(post["n_cases"])):
        H, W, oh, ow, dt, sub, seed0 = (int(v) for v in post[f"post{i}_meta"])
        pp = PostProcessor(foreground_scale=fg, aa_strength=aa)
        for f in range(3):
            raw = torch.from_numpy(synth_depth(seed0 + f, H, W)).to(DT[dt]).to(cuda_device)
            up, low = pp(raw, out_size=(oh, ow), return_lowres=True)
            assert up.dtype == DT[dt] and tuple(up.shape) == (oh, ow)
            e1 = np.abs(low.float().cpu().numpy()[::sub, ::sub] - post[f"post{i}_f{f}_ema"]).max()
            e2 = np.abs(up.float().cpu().numpy()[::sub * 3, ::sub * 3] - post[f"post{i}_f{f}_up"]).max()
            assert e1 <= tol[dt] and e2 <= tol[dt], (i, f, dt, e1, e2)


def test_postprocess_vs_oracle_on_gpu(cuda_device):
    """Same ops through ATen's CUDA kernels (fp16, the dtype the reference computes in on CUDA)."""
    from desktop2stereo_b200.prepost import PostProcessor
    for dt, tol in ((torch.float32, 2e-6), (torch.float16, 2e-3)):
        pp = PostProcessor(foreground_scale=0.05, aa_strength=4.0)
        prev = None
        for f in range(3):
            raw = torch.from_numpy(synth_depth(300 + f, 294, 518)).to(cuda_device).to(dt)
            mine = pp(raw, out_size=(1080, 1920))
            ref = opp.post_process_depth(raw, 0.05, 4.0)
            prev, st = opp.ema(prev, ref)
            ref_up = opp.upsample_depth(st, 1080, 1920)
            err = (mine.float() - ref_up.float()).abs().max().item()
            assert err <= tol, (dt, f, err)


def test_postprocess_edge_cases(cuda_device):
    from desktop2stereo_b200.prepost import PostProcessor
    pp = PostProcessor(foreground_scale=0.0, aa_strength=0.0)   # identity fg-scale, blur disabled (k < 3)
    raw = torch.from_numpy(synth_depth(1, 28, 42)).to(cuda_device)
    out = pp(raw, use_temporal_smooth=False)
    ref = opp.post_process_depth(raw, 0.0, 0.0)
    assert (out - ref).abs().max().item() <= 2e-6
    const = torch.full((28, 42), 3.0, device=cuda_device)       # hi == lo -> denom clamps to 1e-6
    out = pp(const, use_temporal_smooth=False)
    assert torch.equal(out, opp.post_process_depth(const, 0.0, 0.0))
    tiny = torch.rand(2, 5, device=cuda_device)                  # numel <= 10 -> dmin = dmax = 0
    out = pp(tiny, use_temporal_smooth=False)
    assert (out - opp.post_process_depth(tiny, 0.0, 0.0)).abs().max().item() <= 2e-6


def test_postprocess_metric_vs_golden_and_oracle(cuda_device, golden_dir):
    """Metric models (depth.py:837-858): 1/d on d > 0, percentile bounds over the compacted valid values."""
    from desktop2stereo_b200.prepost import PostProcessor
    from oracle.gen_golden import POST_METRIC_CASES, synth_metric_depth
    g = np.load(os.path.join(golden_dir, "post_metric.npz"))
    for (seed, H, W, dt, sub) in POST_METRIC_CASES:
        tdt = getattr(torch, dt)
        raw = torch.from_numpy(synth_metric_depth(seed, H, W)).to(tdt).to(cuda_device)
        pp = PostProcessor(foreground_scale=0.05, aa_strength=4.0, metric=True)
        out = pp(raw, use_temporal_smooth=False)
        tol = 2e-6 if dt == "float32" else 1.6e-2
        assert np.abs(out.float().cpu().numpy()[::sub, ::sub] - g[f"m{seed}"]).max() <= tol, (seed, dt)
        ref = opp.post_process_depth(raw, 0.05, 4.0, metric=True)        # the same ops through ATen's CUDA kernels
        assert (out.float() - ref.float()).abs().max().item() <= tol
    # all-invalid frame: no valid value -> bounds 0/0 -> denom 1e-6 -> everything clamps to 0 (then gamma / fg-scale of 0)
    bad = torch.full((20, 30), -1.0, device=cuda_device)
    out = PostProcessor(foreground_scale=0.05, aa_strength=4.0, metric=True)(bad, use_temporal_smooth=False)
    assert torch.equal(out, opp.post_process_depth(bad, 0.05, 4.0, metric=True))
