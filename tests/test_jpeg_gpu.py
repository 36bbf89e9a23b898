"""d2s_jpeg_encode (device-side baseline JPEG) through the C ABI: the stream is byte-identical to cv2.imencode — the
call MJPEGStreamer makes (reference streamer.py:250-256) — for the same quality and restart interval, and decodes to exactly the
frame the reference's restart-free stream decodes to."""
import numpy as np
import pytest
import torch

from oracle import jpeg as oj

pytestmark = pytest.mark.gpu


def _encode(dev, img, quality, ri, **kw):
    from desktop2stereo_b200.stereo import JpegEncoder
    enc = JpegEncoder(img.shape[0], img.shape[1], dev, quality=quality, restart_interval=ri, **kw)
    return enc.encode_bytes(torch.from_numpy(img).to(dev))


@pytest.mark.parametrize("h,w", [(2, 2), (8, 8), (16, 16), (34, 50), (48, 64), (136, 248), (270, 482), (128, 2064)])
@pytest.mark.parametrize("quality", [20, 90, 100])
def test_jpeg_bytes_equal_cv2_noise(cuda_device, h, w, quality):
    rng = np.random.default_rng(h * 977 + w + quality)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for ri in (1, 3, 8):
        got = _encode(cuda_device, img, quality, ri)
        want = oj.encode_cv2(img, quality, ri)
        assert got == want, (h, w, quality, ri, len(got), len(want))
        assert got == oj.encode_oracle(img, quality, ri)


@pytest.mark.parametrize("h,w,ri", [(1080, 3840, 4), (1080, 1920, 1), (2160, 7680, 8)])
def test_jpeg_full_size_frames(cuda_device, h, w, ri):
    """BASELINE sizes: 1080p Full-SBS (1080 rows: the bottom luma block row is libjpeg's dummy row) and 4K Full-SBS"""
    img = oj.desktop_like(h, w, seed=h + ri)
    got = _encode(cuda_device, img, 90, ri)
    assert got == oj.encode_cv2(img, 90, ri)
    # the reference's own call has no restart markers: same coefficients, so the decoded frames are identical
    assert np.array_equal(oj.decode(got), oj.decode(oj.encode_cv2(img, 90, 0)))


def test_jpeg_extremes_and_row_pitch(cuda_device):
    from desktop2stereo_b200.stereo import JpegEncoder
    img = np.zeros((32, 32, 3), np.uint8)
    img[::2, ::2] = 255                                             # largest AC magnitudes, 0xFF stuffing at quality 100
    assert _encode(cuda_device, img, 100, 1) == oj.encode_cv2(img, 100, 1)
    for v in (0, 128, 255):
        flat = np.full((48, 80, 3), v, np.uint8)
        assert _encode(cuda_device, flat, 90, 2) == oj.encode_cv2(flat, 90, 2)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (66, 94, 3), dtype=np.uint8)
    wide = torch.zeros((66, 101, 3), dtype=torch.uint8, device=cuda_device)      # odd row pitch: the byte-load path
    wide[:, :94] = torch.from_numpy(img).to(cuda_device)
    enc = JpegEncoder(66, 94, cuda_device, quality=75, restart_interval=2)
    assert enc.encode_bytes(wide[:, :94]) == oj.encode_cv2(img, 75, 2)
    # the encoder object is reusable, and a second frame does not see the first one's state
    img2 = rng.integers(0, 256, (66, 94, 3), dtype=np.uint8)
    assert enc.encode_bytes(torch.from_numpy(img2).to(cuda_device)) == oj.encode_cv2(img2, 75, 2)


def test_jpeg_small_capacity_reports_zero(cuda_device):
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.stereo import JpegEncoder
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    enc = JpegEncoder(64, 64, cuda_device, quality=100, restart_interval=1, capacity=2048)
    with pytest.raises(RuntimeError):
        enc.encode_bytes(torch.from_numpy(img).to(cuda_device))
    with pytest.raises(ValueError):
        JpegEncoder(63, 64, cuda_device)
    with pytest.raises(ValueError):
        JpegEncoder(64, 64, cuda_device, restart_interval=0)
    assert _lib.lib().d2s_jpeg_max_bytes(64, 64, 1) > 0


def test_pipeline_jpeg_output(cuda_device):
    """out_format='jpeg': the whole-frame pipeline returns the JPEG of its own u8 frame — byte-identical to cv2.imencode on it —
    for one stream and for several, including a frame that outgrows the adaptive device->host copy size."""
    from desktop2stereo_b200 import depth
    from desktop2stereo_b200.pipeline import StereoPipeline
    from oracle.gen_golden import TINY
    from oracle.ref_harness import make_hf_model
    depth.init(make_hf_model("Small", 3, TINY), device=cuda_device, depth_resolution=126)
    smooth = [np.ascontiguousarray(np.broadcast_to(oj.desktop_like(360, 640, seed=i)[..., [2, 1, 0, 0]], (360, 640, 4))) for i in range(3)]
    noisy = [np.random.default_rng(200 + i).integers(0, 256, (360, 640, 4), dtype=np.uint8) for i in range(3)]
    frames = smooth + noisy + smooth                   # compressible, then 8x larger streams, then compressible again
    p8 = StereoPipeline(depth_slots=3, display_mode="Full-SBS", out_dtype=torch.uint8)
    want = [oj.encode_cv2(r.copy(), 85, 3) for r in p8.run(iter(frames))]
    p8.close()
    pj = StereoPipeline(depth_slots=3, display_mode="Full-SBS", out_dtype=torch.uint8, out_format="jpeg", jpeg_quality=85, jpeg_restart_interval=3)
    got = [bytes(r) for r in pj.run(iter(frames))]
    assert got == want
    assert len(got[4]) > 2 * 65536 + 16 > 2 * len(got[0])       # the noisy frames outgrow the adaptive device->host copy (starts at 128 KB here)
    pj.close()
    # device-resident results carry the same streams
    pj = StereoPipeline(depth_slots=3, display_mode="Full-SBS", out_dtype=torch.uint8, out_format="jpeg", jpeg_quality=85, jpeg_restart_interval=3)
    dev = [bytes(r.cpu().numpy()) for r in pj.run(iter([torch.from_numpy(f).to(cuda_device) for f in frames]), host=False)]
    pj.close()
    assert dev == want
    # several streams per submit
    p8 = StereoPipeline(depth_slots=2, display_mode="Half-SBS", out_dtype=torch.uint8, streams=2)
    pj = StereoPipeline(depth_slots=2, display_mode="Half-SBS", out_dtype=torch.uint8, streams=2, out_format="jpeg")
    pairs = [np.stack([frames[i], frames[i + 3]]) for i in range(3)]
    want2 = [[oj.encode_cv2(r[b].copy(), 90, 4) for b in range(2)] for r in p8.run(iter(pairs))]
    got2 = [[bytes(x) for x in r] for r in pj.run(iter(pairs))]
    p8.close(); pj.close()
    assert got2 == want2


def test_jpeg_matches_committed_golden(cuda_device):
    """the recorded cv2.imencode streams (tests/golden/jpeg.npz), independent of the cv2 build on this box"""
    import os
    from oracle.gen_golden_jpeg import CASES
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg.npz"))
    for name, h, w, content, q, ri in CASES:
        if ri == 0:
            continue                                    # the device encoder always writes restart intervals
        assert _encode(cuda_device, g[name + "/rgb"], q, ri) == g[name + "/jpeg"].tobytes(), name
