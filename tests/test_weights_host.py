"""Host-side weight packing (no GPU): blob sizes follow the documented order, LayerScale folding, the VDA position tables."""
import numpy as np
import torch

from desktop2stereo_b200.synth import TINY_CFG, VDA_ENCODERS, make_hf_model, make_vda_state_dict, param_shapes
from desktop2stereo_b200.weights import config_for_vda, config_from_hf, pack_state_dict, pack_vda_state_dict


def _dav2_count(D, L, c, F):
    n = D * 588 + D + D + (1 + 37 * 37) * D
    n += L * (2 * D + 3 * D * D + 3 * D + D * D + D + 2 * D + 4 * D * D + 4 * D + 4 * D * D + D) + 2 * D
    n += sum(ci * D + ci for ci in c) + c[0] * c[0] * 16 + c[0] + c[1] * c[1] * 4 + c[1] + c[3] * c[3] * 9 + c[3]
    n += sum(F * ci * 9 for ci in c) + 4 * (F * F + F + 4 * (F * F * 9 + F))
    n += (F // 2) * F * 9 + F // 2 + 32 * (F // 2) * 9 + 32 + 32 + 1
    return n


def test_pack_dav2_blob_size_and_layerscale():
    m = make_hf_model("Small", 0, TINY_CFG)
    cfg = config_from_hf(m.config)
    blob = pack_state_dict(m.state_dict(), cfg)
    assert blob.dtype == np.float32 and blob.size == _dav2_count(cfg.hidden, cfg.layers, list(cfg.neck), cfg.fusion)
    sd = m.state_dict()
    D = cfg.hidden
    off = D * 588 + D + D + (1 + 37 * 37) * D + 2 * D + 3 * D * D + 3 * D       # first layer's proj weight
    w = sd["backbone.encoder.layer.0.attention.output.dense.weight"].numpy() * sd["backbone.encoder.layer.0.layer_scale1.lambda1"].numpy()[:, None]
    assert np.array_equal(blob[off:off + D * D].reshape(D, D), w.astype(np.float32))
    assert cfg.temporal == 0 and cfg.pos_interp_offset == 0.0


def test_pack_vda_blob():
    enc = "vits"
    sd = make_vda_state_dict(enc, 3)
    assert [n for n, _ in param_shapes(enc)] == list(sd.keys())
    cfg = config_for_vda(enc)
    e = VDA_ENCODERS[enc]
    assert cfg.temporal == 1 and abs(cfg.pos_interp_offset - 0.1) < 1e-7 and list(cfg.out_indices) == [t + 1 for t in e["taps"]]
    blob = pack_vda_state_dict(sd, cfg)
    base = _dav2_count(cfg.hidden, cfg.layers, list(cfg.neck), cfg.fusion)
    temporal = 0
    for C in (e["out_channels"][2], e["out_channels"][3], e["features"], e["features"]):
        temporal += 2 * C + C * C + C + 2 * (2 * C + 3 * C * C + 32 * 3 * C + C * C + C) + 2 * C + 8 * C * C + 8 * C + 4 * C * C + C + C * C + C
    assert blob.size == base + temporal
    # the position tables are pe @ [to_q; to_k; to_v]^T (the projections are linear and bias-free)
    C = e["out_channels"][2]
    ab = "head.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0."
    qkv = torch.cat([sd[ab + "to_q.weight"], sd[ab + "to_k.weight"], sd[ab + "to_v.weight"]], 0)
    want = (sd[ab + "pos_encoder.pe"][0].double() @ qkv.double().t()).float().numpy()
    off = base + 2 * C + C * C + C + 2 * C + 3 * C * C
    assert np.allclose(blob[off:off + 32 * 3 * C].reshape(32, 3 * C), want, rtol=0, atol=1e-6)


def test_checkpoint_loaders_round_trip(tmp_path):
    """config.json + model.safetensors (HF snapshot layout) and a .pth state_dict (VDA) give the same blob as packing the live
    state_dict."""
    from safetensors.torch import save_file
    from desktop2stereo_b200.checkpoints import load_hf_checkpoint, load_vda_checkpoint
    m = make_hf_model("Small", 1, TINY_CFG)
    d = tmp_path / "hf"
    d.mkdir()
    (d / "config.json").write_text(m.config.to_json_string())
    save_file({k: v.contiguous() for k, v in m.state_dict().items()}, str(d / "model.safetensors"))
    blob, cfg = load_hf_checkpoint(str(d))
    want_cfg = config_from_hf(m.config)
    assert np.array_equal(blob, pack_state_dict(m.state_dict(), want_cfg))
    assert (cfg.hidden, cfg.layers, cfg.heads, list(cfg.neck), cfg.fusion) == (want_cfg.hidden, want_cfg.layers, want_cfg.heads, list(want_cfg.neck), want_cfg.fusion)
    sd = make_vda_state_dict("vits", 5)
    torch.save(sd, str(tmp_path / "video_depth_anything_vits.pth"))
    blob2, cfg2 = load_vda_checkpoint(str(tmp_path / "video_depth_anything_vits.pth"), "vits")
    assert cfg2.temporal == 1 and np.array_equal(blob2, pack_vda_state_dict(sd, config_for_vda("vits")))


def test_hf_snapshot_written_by_save_pretrained(tmp_path):
    """the loader must meet a snapshot as transformers itself writes it (config.json + model.safetensors with the real
    on-disk key names and dtypes), in fp32, fp16 and bf16 (numpy has no bfloat16: the loader reads through torch)."""
    from desktop2stereo_b200.checkpoints import load_hf_checkpoint
    m = make_hf_model("Small", 2, TINY_CFG)
    want_cfg = config_from_hf(m.config)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        d = tmp_path / str(dt).split(".")[-1]
        mm = make_hf_model("Small", 2, TINY_CFG).to(dt)
        mm.save_pretrained(str(d))                      # what `AutoModelForDepthEstimation.from_pretrained` reads (depth.py:1649-1662)
        assert (d / "config.json").exists() and any(f.name.endswith(".safetensors") for f in d.iterdir())
        blob, cfg = load_hf_checkpoint(str(d))
        want = pack_state_dict({k: v.to(dt).float() for k, v in m.state_dict().items()}, want_cfg)
        assert blob.dtype == np.float32 and np.array_equal(blob, want), dt
        assert (cfg.hidden, cfg.layers, cfg.heads, list(cfg.out_indices), list(cfg.neck), cfg.fusion, cfg.metric) == \
               (want_cfg.hidden, want_cfg.layers, want_cfg.heads, list(want_cfg.out_indices), list(want_cfg.neck), want_cfg.fusion, want_cfg.metric)


def test_vda_checkpoint_from_the_reference_module(tmp_path):
    """The .pth the reference loads with strict=True (depth.py:885-901) has exactly the keys of its own VideoDepthAnything module:
    build that module from the UNMODIFIED reference tree, save its state_dict, and check that the loader consumes every tensor
    (names, shapes) and packs the same blob as the synthetic state dict of identical values."""
    import os
    import sys
    import types
    from oracle.ref_harness import REFERENCE_ROOT
    if not os.path.isdir(REFERENCE_ROOT):
        pytest.skip("reference tree not present (GPU box)")
    from desktop2stereo_b200.checkpoints import load_vda_checkpoint
    from desktop2stereo_b200.synth import VDA_ENCODERS, param_shapes
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "easydict" not in sys.modules:      # dpt_temporal.py:19 imports easydict (absent here): attribute-dict stub
        ed = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, **kw):
                super().__init__(**kw)
                self.__dict__ = self
        ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed
    from models.video_depth_anything.vda2_s import VideoDepthAnything
    enc = VDA_ENCODERS["vits"]
    torch.manual_seed(0)
    ref = VideoDepthAnything(encoder="vits", features=enc["features"], out_channels=enc["out_channels"]).eval()
    sd = ref.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(s)) for k, s in param_shapes("vits")]   # names AND shapes, in module order
    path = tmp_path / "video_depth_anything_vits.pth"
    torch.save(sd, str(path))
    blob, cfg = load_vda_checkpoint(str(path), "vits")
    assert cfg.temporal == 1 and np.array_equal(blob, pack_vda_state_dict({k: v.clone() for k, v in sd.items()}, config_for_vda("vits")))
    assert np.isfinite(blob).all()
