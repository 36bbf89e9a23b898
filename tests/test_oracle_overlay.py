"""The overlay oracle (oracle/overlay.py) against the reference's own overlay_fps output (tests/golden/overlay.npz,
written by oracle/gen_golden.py from the unmodified reference): bit-exact, including the every-10th-call text cache."""
import os

import numpy as np
import pytest

from oracle.gen_golden import OVERLAY_CASES, overlay_rgb
from oracle.overlay import OverlayOracle

NP_DT = {"float32": np.float32, "float16": np.float16}


@pytest.mark.parametrize("case", OVERLAY_CASES, ids=lambda c: f"ov{c[0]}_{c[1]}x{c[2]}_{c[3]}")
def test_overlay_oracle_matches_reference(golden_dir, case):
    seed, H, W, dt, fps_seq = case
    gold = np.load(os.path.join(golden_dir, "overlay.npz"))[f"ov{seed}"]
    if dt == "bfloat16":
        import torch
        rgb = overlay_rgb(seed, H, W)          # integers 0..255 are exact in bf16; the arithmetic is exact for a 0/1 mask
    else:
        rgb = overlay_rgb(seed, H, W).astype(NP_DT[dt])
    o = OverlayOracle()
    ch, cw = min(H, 64), min(W, 420)
    for i, fps in enumerate(fps_seq):
        out = o(rgb, fps)
        assert np.array_equal(out[:, :ch, :cw].astype(np.float32).astype(np.uint8), gold[i]), (seed, i)
        assert np.array_equal(out[:, ch:], rgb[:, ch:]) and np.array_equal(out[:, :, cw:], rgb[:, :, cw:])
