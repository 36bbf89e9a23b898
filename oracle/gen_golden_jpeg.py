"""Generates tests/golden/jpeg.npz: seeded frames and the JPEG streams cv2.imencode (OpenCV's bundled libjpeg-turbo; the call the
reference's MJPEGStreamer makes, streamer.py:250-256) writes for them, so the JPEG parity tests have a fixture that does not
depend on the cv2 build present at test time.  Run in the build container: python -m oracle.gen_golden_jpeg"""
import os

import numpy as np

from . import jpeg as oj

CASES = [  # (name, h, w, content, quality, restart interval)
    ("noise_48x64_q90_ri3", 48, 64, "noise", 90, 3),
    ("noise_34x50_q100_ri1", 34, 50, "noise", 100, 1),
    ("noise_136x248_q20_ri8", 136, 248, "noise", 20, 8),
    ("desktop_120x216_q90_ri4", 120, 216, "desktop", 90, 4),
    ("desktop_120x216_q75_ri0", 120, 216, "desktop", 75, 0),
]


def frame(h, w, content, seed):
    if content == "noise":
        return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    return oj.desktop_like(h, w, seed)


def main():
    import cv2
    out = {"cv2_version": np.array(cv2.__version__), "jpeg_library": np.array([l.strip() for l in cv2.getBuildInformation().splitlines() if "JPEG:" in l][0])}
    for i, (name, h, w, content, q, ri) in enumerate(CASES):
        img = frame(h, w, content, 1000 + i)
        out[name + "/rgb"] = img
        out[name + "/jpeg"] = np.frombuffer(oj.encode_cv2(img, q, ri), np.uint8)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "jpeg.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
