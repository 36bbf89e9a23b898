"""ORACLE (test infrastructure) — baseline JPEG of an RGB u8 frame, as cv2.imencode(".jpg") produces it for the reference's
MJPEGStreamer (reference streamer.py:250-256).  Two checkers:
    encode_cv2     the reference's own call (OpenCV's bundled libjpeg-turbo) with the same quality and restart interval — the pin;
    encode_oracle  oracle/jpeg_oracle.c, the stage-by-stage C restatement (byte-identical to encode_cv2, tests/test_oracle_jpeg.py)."""
import ctypes

import numpy as np

from . import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.d2s_oracle_jpeg_encode.restype = ctypes.c_size_t
        _lib.d2s_oracle_jpeg_encode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    return _lib


def encode_oracle(rgb: np.ndarray, quality: int = 90, restart_interval: int = 0) -> bytes:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    out = np.empty(h * w * 4 + 4096, np.uint8)
    n = _load().d2s_oracle_jpeg_encode(rgb.ctypes.data, h, w, quality, restart_interval, out.ctypes.data, out.size)
    assert n > 0, "oracle jpeg: output buffer too small or bad size"
    return out[:n].tobytes()


def encode_cv2(rgb: np.ndarray, quality: int = 90, restart_interval: int = 0) -> bytes:
    import cv2
    bgr = np.ascontiguousarray(rgb[..., ::-1])                       # streamer.py:249
    params = [cv2.IMWRITE_JPEG_QUALITY, int(quality)]
    if restart_interval:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(restart_interval)]
    ok, buf = cv2.imencode(".jpg", bgr, params)
    assert ok
    return buf.tobytes()


def decode(jpeg: bytes) -> np.ndarray:
    import cv2
    return cv2.imdecode(np.frombuffer(jpeg, np.uint8), cv2.IMREAD_COLOR)[..., ::-1]


def desktop_like(h: int, w: int, seed: int = 0) -> np.ndarray:
    """A frame with the statistics of desktop content (flat panels, gradients, sharp text-like edges, one noisy 'video' window):
    JPEG size depends on content, so compact-output measurements name the content they used."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([40 + 60 * xx / w, 50 + 80 * yy / h, 90 + 40 * (xx + yy) / (h + w)], -1)
    for _ in range(24):                                              # windows
        y0, x0 = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
        hh, ww = int(rng.integers(8, max(9, h // 3))), int(rng.integers(8, max(9, w // 3)))
        img[y0:y0 + hh, x0:x0 + ww] = rng.integers(0, 256, 3)
    ty, tx = h // 8, w // 8                                          # "text": 1-px-wide random strokes on a light panel
    panel = img[ty:ty + h // 4, tx:tx + w // 3]
    panel[:] = 235
    mask = rng.random(panel.shape[:2]) < 0.18
    mask[::3] = False
    panel[mask] = 20
    vy, vx = h // 2, w // 2                                          # "video": smooth texture + mild noise
    vh, vw = h // 3, w // 3
    tex = 128 + 60 * np.sin(xx[vy:vy + vh, vx:vx + vw] / 9.0) * np.cos(yy[vy:vy + vh, vx:vx + vw] / 7.0)
    img[vy:vy + vh, vx:vx + vw] = tex[..., None] + rng.normal(0, 6, (tex.shape[0], tex.shape[1], 3))
    return np.clip(img, 0, 255).astype(np.uint8)
