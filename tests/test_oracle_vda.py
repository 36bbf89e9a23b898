"""The streaming Video-Depth-Anything oracle (oracle/vda.py) against the reference's own module (tests/golden/vda.npz, written
by oracle/gen_golden.py from models/video_depth_anything/vda2_s.py in fp32): 40 frames, so the 32-frame cache wraps."""
import os

import numpy as np
import torch

from oracle import vda
from oracle.gen_golden import VDA_CASE, vda_frames


def test_vda_oracle_matches_reference(golden_dir):
    c = VDA_CASE
    g = np.load(os.path.join(golden_dir, "vda.npz"))
    o = vda.StreamingVDA(vda.make_state_dict(c["encoder"], c["seed"]), c["encoder"])
    frames = vda_frames(c["seed"], c["frames"], c["H"], c["W"])
    worst = 0.0
    for t in range(c["frames"]):
        d = o(torch.from_numpy(frames[t]))[0, 0].numpy()
        if t in c["keep"]:
            ref = g[f"depth{t}"]
            worst = max(worst, float(np.abs(d - ref).max() / np.abs(ref).max()))
    assert worst <= 2e-5, worst            # fp32 vs fp32: summation-order noise only


def test_vda_param_table_is_complete():
    for enc in vda.ENCODERS:
        names = [n for n, _ in vda.param_shapes(enc)]
        assert len(names) == len(set(names))
    assert sum(int(np.prod(s)) for _, s in vda.param_shapes("vits")) == 29080193 - 0   # the reference's vits checkpoint size
