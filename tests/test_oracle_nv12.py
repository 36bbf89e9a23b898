"""oracle/nv12.py (libjpeg's RGB->YCbCr + h2v2 downsample, the first stages of the cv2.imencode the reference runs in
streamer.py:250-256) pinned on OpenCV's own BT.601 full-range conversion, plus known answers."""
import numpy as np

from oracle import nv12


def test_known_answers():
    for rgb, (y, cb, cr) in {(0, 0, 0): (0, 128, 128), (255, 255, 255): (255, 128, 128), (255, 0, 0): (76, 85, 255),
                             (0, 255, 0): (150, 44, 21), (0, 0, 255): (29, 255, 107), (128, 128, 128): (128, 128, 128)}.items():
        img = np.tile(np.array(rgb, np.uint8), (2, 2, 1))
        out = nv12.rgb_to_nv12(img)
        assert out.shape == (3, 2)
        assert (out[:2] == y).all() and out[2, 0] == cb and out[2, 1] == cr, (rgb, out)


def test_matches_opencv_ycrcb():
    import cv2
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    out = nv12.rgb_to_nv12(img)
    ycc = cv2.cvtColor(img, cv2.COLOR_RGB2YCrCb)            # same matrix (BT.601 full range), OpenCV's own fixed-point rounding
    assert np.abs(out[:64].astype(int) - ycc[..., 0].astype(int)).max() <= 1
    cb_full, cr_full = ycc[..., 2].astype(np.float64), ycc[..., 1].astype(np.float64)
    down = lambda c: (c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]) / 4
    uv = out[64:].reshape(32, 48, 2).astype(np.float64)
    assert np.abs(uv[..., 0] - down(cb_full)).max() <= 1.5 and np.abs(uv[..., 1] - down(cr_full)).max() <= 1.5
