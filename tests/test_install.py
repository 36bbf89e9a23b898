"""install(): patching the reference `depth` module in place (SURVEY §8b import-time contract)."""
import os
import sys
import textwrap
import types

import numpy as np
import pytest
import torch

from oracle.ref_harness import REFERENCE_ROOT


class _StubEngine:
    """Stands in for B200Engine where there is no GPU: install() only stores it in the slot."""
    backend_name = "B200"

    def __init__(self):
        self.device = torch.device("cpu")
        self.cfg = types.SimpleNamespace(temporal=0)

    def __call__(self, t):
        raise AssertionError("not called on CPU")


# a module shaped like the reference's depth.py: the wrapper singleton is created in the MIDDLE (depth.py:1784), the functions
# main.py imports are defined BELOW it (:1897, :2122, :2186)
_REF_LIKE = textwrap.dedent('''
    DEPTH_RESOLUTION = 126
    FP16 = True
    FOREGROUND_SCALE = 0.05
    AA_STRENGTH = 4.0
    class DepthModelWrapper:
        def __init__(self):
            self.model, self.backend = MODEL, "PyTorch"
        def __call__(self, t):
            return self.model(t)
    def process(img, target_height): raise RuntimeError("reference process")
    model_wraper = DepthModelWrapper()
    {mid}
    def predict_depth(image_rgb, return_tuple=False, use_temporal_smooth=True): raise RuntimeError("reference predict_depth")
    def make_sbs_core(rgb, depth, **kw): raise RuntimeError("reference make_sbs_core")
    def make_sbs(rgb_c, depth, **kw): raise RuntimeError("reference make_sbs")
    {end}
''')


def _ref_like(name, model, mid="", end=""):
    m = types.ModuleType(name)
    m.MODEL = model
    sys.modules[name] = m
    try:
        exec(compile(_REF_LIKE.format(mid=mid, end=end), name + ".py", "exec"), m.__dict__)
    finally:
        sys.modules.pop(name, None)
    return m


def test_install_at_end_of_module_patches_every_name():
    import desktop2stereo_b200.depth as b200
    eng = _StubEngine()
    m = types.ModuleType("ref_like_end")
    m.MODEL, m.ENGINE, m.b200 = object(), eng, b200
    sys.modules["ref_like_end"] = m
    try:
        exec(compile(_REF_LIKE.format(mid="", end="b200.install(__import__('sys').modules[__name__], engine=ENGINE, device='cpu')"),
                     "ref_like_end.py", "exec"), m.__dict__)
    finally:
        sys.modules.pop("ref_like_end", None)
    for n in ("process", "predict_depth", "make_sbs", "make_sbs_core"):
        assert getattr(m, n) is getattr(b200, n), n
    assert m.model_wraper.model is eng and m.model_wraper.backend == "B200"
    assert b200.settings.depth_resolution == 126 and b200.settings.aa_strength == 4.0    # taken from the module's constants


def test_install_in_the_middle_is_refused():
    """Called where the reference creates model_wraper (depth.py:1784) the functions below would overwrite the patch: install()
    sees that they do not exist yet and raises instead of patching half of the names."""
    import desktop2stereo_b200.depth as b200
    from desktop2stereo_b200 import _lib
    m = types.ModuleType("ref_like_mid")
    m.MODEL, m.ENGINE, m.b200 = object(), _StubEngine(), b200
    sys.modules["ref_like_mid"] = m
    try:
        with pytest.raises(_lib.D2SError, match="executed completely"):
            exec(compile(_REF_LIKE.format(mid="b200.install(__import__('sys').modules[__name__], engine=ENGINE, device='cpu')", end=""),
                         "ref_like_mid.py", "exec"), m.__dict__)
    finally:
        sys.modules.pop("ref_like_mid", None)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference tree not present (GPU box)")
def test_install_on_the_unmodified_reference_module():
    """The real reference depth.py, imported through the harness: after install() the names main.py imports (main.py:44, :1321)
    are the B200 ones, the engine sits in the wrapper slot, and the settings came from the module's own constants."""
    import desktop2stereo_b200.depth as b200
    from oracle.ref_harness import load_reference
    ref = load_reference("Small", depth_resolution=336)
    eng = _StubEngine()
    before = ref.predict_depth
    b200.install(ref, engine=eng, device="cpu")
    assert ref.predict_depth is b200.predict_depth and ref.predict_depth is not before
    assert ref.process is b200.process and ref.make_sbs is b200.make_sbs and ref.make_sbs_core is b200.make_sbs_core
    assert ref.model_wraper.model is eng and ref.model_wraper.backend == "B200"
    assert b200.settings.depth_resolution == 336 and b200.settings.fp16 == ref.FP16
    assert abs(b200.settings.foreground_scale - ref.FOREGROUND_SCALE) < 1e-12 and b200.settings.aa_strength == ref.AA_STRENGTH
    # main.py's import form picks up the patched functions
    ns = {}
    sys.modules["depth"] = ref
    try:
        exec("from depth import process, predict_depth\nfrom depth import make_sbs", ns)
    finally:
        sys.modules.pop("depth", None)
    assert ns["predict_depth"] is b200.predict_depth and ns["make_sbs"] is b200.make_sbs


@pytest.mark.gpu
def test_install_end_to_end_through_patched_names(cuda_device):
    """A reference-shaped module patched in place: its own names run the B200 path end to end, and the engine slot is called
    the way DepthModelWrapper.__call__ calls it (depth.py:1781)."""
    import desktop2stereo_b200.depth as b200
    from oracle.gen_golden import TINY, synth_frame
    from oracle.ref_harness import make_hf_model
    model = make_hf_model("Small", 2, TINY)
    m = _ref_like("ref_like_gpu", model)
    b200.install(m, device=cuda_device, depth_resolution=70)      # no engine given: packs m.model_wraper.model (the HF module)
    assert m.model_wraper.backend == "B200"
    frame = synth_frame(0, 90, 160, 4)
    rgb = m.process(frame, 90)
    d = m.predict_depth(rgb)
    sbs = m.make_sbs(rgb, d, display_mode="Half-SBS")
    assert sbs.shape == (90, 160, 3) and sbs.dtype == np.float32 and np.isfinite(sbs).all()
    x = torch.zeros(1, 3, 70, 126, device=cuda_device)
    assert tuple(m.model_wraper(x).shape) == (1, 70, 126)


def test_jpeg_frame_parsing_host_logic():
    """pipeline._jpeg_streams: d2s_pipe_jpeg_frame buffers (u32 size, 12 reserved bytes, stream) -> the streams they hold"""
    import numpy as np
    from desktop2stereo_b200.pipeline import _jpeg_streams
    buf = np.zeros((2, 64), np.uint8)
    for b, payload in enumerate((b"\xff\xd8abc\xff\xd9", b"\xff\xd8\xff\xd9")):
        buf[b, :4] = np.frombuffer(np.uint32(len(payload)).tobytes(), np.uint8)
        buf[b, 16:16 + len(payload)] = np.frombuffer(payload, np.uint8)
    one = _jpeg_streams(buf[0], 1)
    assert bytes(one) == b"\xff\xd8abc\xff\xd9"
    two = _jpeg_streams(buf, 2)
    assert [bytes(x) for x in two] == [b"\xff\xd8abc\xff\xd9", b"\xff\xd8\xff\xd9"]
