"""Pin oracle/prepost.py on goldens produced by the unmodified reference (tests/golden/pre.npz, post.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import prepost as opp
from oracle.gen_golden import synth_depth, synth_frame

DT = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}


@pytest.fixture(scope="module")
def pre(golden_dir):
    return np.load(os.path.join(golden_dir, "pre.npz"))


@pytest.fixture(scope="module")
def post(golden_dir):
    return np.load(os.path.join(golden_dir, "post.npz"))


def test_model_input_shapes_bit_exact(pre):
    for h, w, target, nh, nw in pre["shapes"]:
        assert opp.model_input_shape(int(h), int(w), int(target)) == (int(nh), int(nw)), (h, w, target)
    assert opp.model_input_shape(1080, 1920) == (294, 518)
    assert opp.model_input_shape(2160, 3840) == (294, 518)
    assert opp.model_input_shape(518, 518) == (518, 518)


def test_process_cuda_branch(pre):
    for seed in (0, 1, 2):
        _, h, w, ch, th = (int(v) for v in pre[f"proc{seed}_meta"])
        frame = torch.from_numpy(synth_frame(seed, h, w, ch))
        for name, dt in (("float32", torch.float32), ("float16", torch.float16)):
            key = f"proc{seed}_{name}"
            if key not in pre.files:
                continue
            out = opp.process_cuda_branch(frame, th, dt).float().numpy()
            assert np.array_equal(out, pre[key].astype(np.float32)), key


def test_resize_and_normalise(pre):
    for seed in (10, 11, 12, 13, 14):
        _, h, w, target = (int(v) for v in pre[f"pre{seed}_meta"])
        frame = synth_frame(seed, h, w, 3)
        t = torch.from_numpy(frame[..., ::-1].copy()).permute(2, 0, 1).unsqueeze(0)
        r = opp.resize_patch_aligned(t, target, 14)
        assert np.array_equal(r.float().numpy()[0], pre[f"pre{seed}_resized"]), seed
        x = opp.normalise_input(r)
        assert np.array_equal(x.numpy()[0], pre[f"pre{seed}_input"]), seed


def test_post_process_ema_upsample(post):
    fg, aa = (float(v) for v in post["fg_aa"])
    tol = {0: 3e-7, 1: 1e-3, 2: 8e-3}   # <= 1 ulp of the compute dtype on [0,1]
    nt = torch.get_num_threads()
    torch.set_num_threads(1)            # the reference pins one thread (depth.py:19)
    try:
        _check_post(post, fg, aa, tol)
    finally:
        torch.set_num_threads(nt)


def _check_post(post, fg, aa, tol):
    for i in range(int(post["n_cases"])):
        H, W, oh, ow, dt, sub, seed0 = (int(v) for v in post[f"post{i}_meta"])
        prev = None
        for f in range(3):
            raw = torch.from_numpy(synth_depth(seed0 + f, H, W)).to(DT[dt])
            pp = opp.post_process_depth(raw, fg, aa)
            prev, st = opp.ema(prev, pp)
            up = opp.upsample_depth(st, oh, ow)
            assert np.allclose(pp.float().numpy()[::sub, ::sub], post[f"post{i}_f{f}_pp"], rtol=0, atol=tol[dt]), (i, f, "pp")
            assert np.allclose(st.float().numpy()[::sub, ::sub], post[f"post{i}_f{f}_ema"], rtol=0, atol=tol[dt]), (i, f, "ema")
            # ATen's CPU bilinear kernel differs by 1 ulp between its vector body and scalar tail, which move with
            # the thread count (the reference pins 1 thread, depth.py:19): floating-point stage, tolerance 1 ulp.
            assert np.allclose(up.float().numpy()[::sub * 3, ::sub * 3], post[f"post{i}_f{f}_up"], rtol=0, atol=tol[dt]), (i, f, "up")


def test_post_metric_oracle_matches_reference(golden_dir):
    """Metric models: 1/d on the valid mask, percentile bounds over the compacted valid values (depth.py:837-858)."""
    import os
    import numpy as np
    import torch
    from oracle import prepost as opp
    from oracle.gen_golden import POST_METRIC_CASES, synth_metric_depth
    g = np.load(os.path.join(golden_dir, "post_metric.npz"))
    for (seed, H, W, dt, sub) in POST_METRIC_CASES:
        raw = torch.from_numpy(synth_metric_depth(seed, H, W)).to(getattr(torch, dt))
        got = opp.post_process_depth(raw, 0.05, 4.0, metric=True).float().numpy()[::sub, ::sub]
        tol = 2e-6 if dt == "float32" else 8e-3      # 1 bf16 ulp at 1.0
        assert np.abs(got - g[f"m{seed}"]).max() <= tol, (seed, np.abs(got - g[f"m{seed}"]).max())
