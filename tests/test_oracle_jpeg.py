"""The JPEG oracle (oracle/jpeg_oracle.c) pinned byte for byte on cv2.imencode — the call the reference's MJPEGStreamer makes on
every frame (reference streamer.py:250-256).  cv2 (OpenCV's bundled libjpeg-turbo) is the reference's own dependency and is present
both here and on the GPU box."""
import numpy as np
import pytest

from oracle import jpeg as oj

cv2 = pytest.importorskip("cv2")

SIZES = [(2, 2), (8, 8), (16, 16), (34, 50), (48, 64), (270, 482), (136, 248)]


@pytest.mark.parametrize("h,w", SIZES)
@pytest.mark.parametrize("quality", [20, 50, 90, 100])
def test_oracle_matches_cv2_noise(h, w, quality):
    rng = np.random.default_rng(h * 1000 + w + quality)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for ri in (0, 1, 3, 8):
        assert oj.encode_oracle(img, quality, ri) == oj.encode_cv2(img, quality, ri), (h, w, quality, ri)


def test_oracle_matches_cv2_desktop_like_1080():
    img = oj.desktop_like(1080, 1920, seed=3)          # 1080 = 67.5 MCU rows: the bottom luma block row is libjpeg's dummy row
    for ri in (0, 4):
        assert oj.encode_oracle(img, 90, ri) == oj.encode_cv2(img, 90, ri)


def test_extreme_values_and_flat():
    for v in (0, 255, 128):
        img = np.full((32, 48, 3), v, np.uint8)
        assert oj.encode_oracle(img, 90, 2) == oj.encode_cv2(img, 90, 2)
    img = np.zeros((32, 32, 3), np.uint8)
    img[::2, ::2] = 255                                 # largest AC magnitudes; exercises 0xFF stuffing at quality 100
    assert oj.encode_oracle(img, 100, 1) == oj.encode_cv2(img, 100, 1)


def test_restart_markers_do_not_change_pixels():
    """the device encoder always writes restart intervals; the reference's call writes none.  Same coefficients -> same decoded frame."""
    img = oj.desktop_like(144, 256, seed=1)
    a = oj.decode(oj.encode_cv2(img, 90, 0))
    b = oj.decode(oj.encode_oracle(img, 90, 4))
    assert np.array_equal(a, b)

