"""Pin the warp oracle (oracle/warp_oracle.c) on goldens produced by the unmodified reference
(oracle/gen_golden.py -> tests/golden/warp.npz).  Bit-exact: the oracle's strict-fp32 restatement
reproduces make_sbs_core's CPU output to the last bit in every mode/dtype/branch."""
import os

import numpy as np
import pytest

from oracle import warp

MODES = ["Full-SBS", "Half-SBS", "Full-TAB", "Half-TAB"]
DTS = ["float32", "float16", "bfloat16"]


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "warp.npz"))


def test_golden_cases_bit_exact(g):
    n = int(g["n_cases"])
    assert n >= 100
    for i in range(n):
        k = f"c{i:03d}"
        h, w, mode, fill, dt, gather = (int(v) for v in g[k + "_meta"])
        ipd, ratio, conv = (float(v) for v in g[k + "_par"])
        out = warp.make_sbs_core_oracle(g[k + "_rgb"].astype(np.float32), g[k + "_depth"], ipd, ratio, MODES[mode],
                                        bool(fill), conv, depth_dtype=DTS[dt], gather=bool(gather))
        ref = g[k + "_out"]
        assert out.shape == ref.shape, (k, out.shape, ref.shape)
        assert np.array_equal(out, ref), (k, MODES[mode], fill, DTS[dt], gather, np.abs(out - ref).max())


def test_known_answer_indices(g):
    """SURVEY §8c KAT (i): H=4, W=16, depth = linspace(0,1,16): gather coords L=[0,0,1,..,14], R=[0..15];
    grid_sample floor(ix) = [0,0,1,..,14]."""
    rgb, dep = g["kat_rgb"], g["kat_depth"]
    out, il, ir = warp.make_sbs_core_oracle(rgb, dep, display_mode="Full-SBS", gather=True, return_indices=True)
    assert np.array_equal(out, g["kat_out_gather"])
    assert il[0].tolist() == [0, 0] + list(range(1, 15))
    assert ir[0].tolist() == list(range(16))
    out, il, ir = warp.make_sbs_core_oracle(rgb, dep, display_mode="Full-SBS", return_indices=True)
    assert np.array_equal(out, g["kat_out_bilinear"])
    assert il[0].tolist() == [0, 0] + list(range(1, 15))


def test_torch_restatement_matches_oracle():
    """The torch restatement (used on the GPU box to drive ATen's own CUDA kernels) agrees with the C oracle on CPU."""
    import torch
    rng = np.random.default_rng(7)
    for (h, w) in [(17, 33), (40, 64)]:
        for mode in MODES:
            for fill in (False, True):
                rgb = rng.integers(0, 256, (3, h, w)).astype(np.float32)
                dep = rng.random((h, w)).astype(np.float32)
                for gather in (False, True):
                    a = warp.make_sbs_core_oracle(rgb, dep, 0.064, 4.0, mode, fill, 0.25, gather=gather)
                    b = warp.make_sbs_core_torch(torch.from_numpy(rgb), torch.from_numpy(dep), 0.064, 4.0, mode, fill, 0.25,
                                                 gather=gather).numpy()
                    assert np.array_equal(a, b), (h, w, mode, fill, gather)


def test_linspace_matches_torch():
    import torch
    for n in (2, 3, 16, 518, 1080, 1920, 2160, 3840):
        assert np.array_equal(warp.linspace(n), torch.linspace(-1, 1, n).numpy()), n


def test_edge_shapes():
    rng = np.random.default_rng(3)
    # h == 1, tiny w, huge shifts that reflect at both borders
    for (h, w, ratio) in [(1, 2, 2.0), (2, 3, 50.0), (3, 5, 400.0)]:
        rgb = rng.integers(0, 256, (3, h, w)).astype(np.float32)
        dep = rng.random((h, w)).astype(np.float32)
        import torch
        a = warp.make_sbs_core_oracle(rgb, dep, 0.064, ratio, "Full-SBS", False, 0.0)
        b = warp.make_sbs_core_torch(torch.from_numpy(rgb), torch.from_numpy(dep), 0.064, ratio, "Full-SBS", False, 0.0).numpy()
        assert np.array_equal(a, b), (h, w, ratio)
