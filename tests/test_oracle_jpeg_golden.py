"""oracle/jpeg_oracle.c against the committed cv2.imencode streams (tests/golden/jpeg.npz, written by oracle/gen_golden_jpeg.py in the
build container): this pin does not need cv2 at test time."""
import numpy as np

from oracle import jpeg as oj


def test_oracle_matches_committed_golden():
    """tests/golden/jpeg.npz: cv2.imencode streams recorded in the build container (oracle/gen_golden_jpeg.py)"""
    import os
    from oracle.gen_golden_jpeg import CASES
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg.npz"))
    for name, h, w, content, q, ri in CASES:
        assert oj.encode_oracle(g[name + "/rgb"], q, ri) == g[name + "/jpeg"].tobytes(), name
