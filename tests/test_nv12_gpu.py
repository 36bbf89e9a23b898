"""d2s_rgb_to_nv12 (libjpeg's colour conversion + 4:2:0 downsample on the device) through the C ABI: bit-exact against
oracle/nv12.py; and the pipeline's NV12 output == NV12 of its own u8 frame."""
import numpy as np
import pytest
import torch

from oracle import nv12
from oracle.gen_golden import TINY, synth_frame
from oracle.ref_harness import make_hf_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w", [(2, 2), (6, 10), (270, 960), (1080, 3840), (34, 62)])
def test_nv12_bit_exact(cuda_device, h, w):
    from desktop2stereo_b200.stereo import rgb_to_nv12
    rng = np.random.default_rng(h * w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = rgb_to_nv12(torch.from_numpy(img).to(cuda_device)).cpu().numpy()
    assert got.shape == (h * 3 // 2, w) and np.array_equal(got, nv12.rgb_to_nv12(img))
    # a row-padded view (pitch > 3 * w) gives the same frame
    wide = torch.zeros((h, w + 5, 3), dtype=torch.uint8, device=cuda_device)
    wide[:, :w] = torch.from_numpy(img).to(cuda_device)
    assert np.array_equal(rgb_to_nv12(wide[:, :w]).cpu().numpy(), got)


def test_nv12_rejects_odd_sizes(cuda_device):
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.stereo import rgb_to_nv12
    with pytest.raises(_lib.D2SError):
        rgb_to_nv12(torch.zeros((5, 8, 3), dtype=torch.uint8, device=cuda_device))


def test_pipeline_nv12_output(cuda_device):
    from desktop2stereo_b200 import depth
    from desktop2stereo_b200.pipeline import StereoPipeline
    depth.init(make_hf_model("Small", 3, TINY), device=cuda_device, depth_resolution=126)
    frames = [synth_frame(200 + i, 180, 320, 4) for i in range(6)]
    p8 = StereoPipeline(depth_slots=3, display_mode="Full-SBS", out_dtype=torch.uint8)
    want = [nv12.rgb_to_nv12(r.copy()) for r in p8.run(iter(frames))]
    p8.close()
    pn = StereoPipeline(depth_slots=3, display_mode="Full-SBS", out_dtype=torch.uint8, out_format="nv12")
    got = [r.copy() for r in pn.run(iter(frames))]
    pn.close()
    for g, w_ in zip(got, want):
        assert g.shape == (270, 640) and g.dtype == np.uint8 and np.array_equal(g, w_)
