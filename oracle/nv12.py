"""ORACLE (test infrastructure) — RGB u8 -> NV12 as libjpeg computes its first two encoder stages, which is what
cv2.imencode(".jpg") runs on the frame the reference hands to MJPEGStreamer (reference streamer.py:250-256):
    rgb_ycc_convert   jccolor.c: JFIF full-range BT.601 in 16-bit fixed point
                      Y  = ( 19595 R + 38470 G +  7471 B + 32768) >> 16
                      Cb = (-11059 R - 21709 G + 32768 B + (128 << 16) + 32767) >> 16
                      Cr = ( 32768 R - 27439 G -  5329 B + (128 << 16) + 32767) >> 16
    h2v2_downsample   jcsample.c: mean of each 2x2 block with the alternating rounding bias 1, 2, 1, 2, ... along a row
libjpeg (IJG / libjpeg-turbo, the JPEG library OpenCV links) is not vendored in /root/reference; its published algorithm is
restated here.  Pinned by tests/test_oracle_nv12.py against cv2's own RGB->YCrCb conversion (8-bit, same BT.601 full-range matrix)."""
import numpy as np


def rgb_to_nv12(rgb: np.ndarray) -> np.ndarray:
    """rgb [h, w, 3] u8 (h, w even) -> nv12 [h * 3 // 2, w] u8: h rows of Y, then h/2 rows of interleaved Cb Cr"""
    h, w, _ = rgb.shape
    assert h % 2 == 0 and w % 2 == 0
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16
    bias = np.tile(np.array([1, 2], np.int64), w // 4 + 1)[: w // 2][None, :]

    def down(c):
        return (c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2] + bias) >> 2
    uv = np.stack([down(cb), down(cr)], -1).reshape(h // 2, w)
    return np.concatenate([y, uv], 0).astype(np.uint8)
