"""Import the UNMODIFIED reference `depth.py` on CPU (test infrastructure only).

This file is part of the ORACLE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may use it.  It never ships on the product path.

It follows the recipe recorded in SURVEY.md §8(c):
  1. a scratch CWD holding a `settings.yaml` with the keys `utils.py:819-907` reads,
  2. `sys.path.insert(0, <reference>)`,
  3. `transformers.AutoModelForDepthEstimation.from_pretrained` replaced by a function that
     returns a seeded random-init DepthAnythingForDepthEstimation (no network, no weights on disk),
  4. `import depth`.

`/root/reference` exists only in the build container.  On the GPU box this module is unusable
and `load_reference()` raises; everything that travels uses the committed goldens + `oracle/*.py`.
"""
from __future__ import annotations

import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("D2S_REFERENCE_ROOT", "/root/reference")

# name -> (hidden, layers, heads, out_indices, neck_hidden_sizes, fusion_hidden_size)
# depth.py:889-893 (VDA table) and the HF `-hf` checkpoints' config.json.
DA_V2_VARIANTS = {
    "Small": (384, 12, 6, [3, 6, 9, 12], [48, 96, 192, 384], 64),
    "Base": (768, 12, 12, [3, 6, 9, 12], [96, 192, 384, 768], 128),
    "Large": (1024, 24, 16, [5, 12, 18, 24], [256, 512, 1024, 1024], 256),
}


def make_hf_model(variant: str = "Small", seed: int = 0, tiny: dict | None = None):
    """Seeded random-init HF DepthAnythingForDepthEstimation (fp32, eval).

    `tiny` overrides (hidden, layers, heads, out_indices, neck, fusion) for small golden cases.
    Parameters that HF initialises to constants (LayerScale=1, biases=0, LN) are re-drawn so that
    every term of the forward pass is exercised by parity tests.
    """
    import torch
    from transformers import DepthAnythingConfig, DepthAnythingForDepthEstimation, Dinov2Config

    hidden, layers, heads, out_idx, neck, fusion = DA_V2_VARIANTS[variant] if tiny is None else (
        tiny["hidden"], tiny["layers"], tiny["heads"], tiny["out_indices"], tiny["neck"], tiny["fusion"])
    bcfg = Dinov2Config(
        hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
        image_size=518, patch_size=14, out_indices=out_idx,
        apply_layernorm=True, reshape_hidden_states=False,
    )
    cfg = DepthAnythingConfig(
        backbone_config=bcfg, reassemble_hidden_size=hidden, patch_size=14,
        neck_hidden_sizes=neck, fusion_hidden_size=fusion, head_hidden_size=32,
        reassemble_factors=[4, 2, 1, 0.5], head_in_index=-1,
        depth_estimation_type="relative",
    )
    torch.manual_seed(seed)
    model = DepthAnythingForDepthEstimation(cfg).eval()
    randomize_constant_params(model, seed + 1)
    return model


def randomize_constant_params(model, seed: int):
    """Give biases / LayerNorm / LayerScale non-trivial seeded values (deterministic by name order)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if name.endswith("weight") and p.dim() >= 2:
                # variance-preserving re-draw: HF's default trunc-normal(0.02) init makes every activation collapse
                # towards 0 and the final ReLU output identically 0, which would make parity tests vacuous
                fan_in = p[0].numel() if "resize" not in name or p.dim() != 4 or "layers.3" in name else p.shape[0] * p[0, 0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / fan_in ** 0.5))
                continue
            if name == "head.conv3.bias":
                p.fill_(0.5)
                continue
            if name.endswith("lambda1"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif name.endswith("cls_token") or name.endswith("position_embeddings"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


def load_reference(variant: str = "Small", depth_resolution: int = 518, fp16: bool = False,
                   seed: int = 0, display_mode: str = "Half-SBS", tiny: dict | None = None,
                   extra_settings: dict | None = None):
    """Return the reference's `depth` module, imported fresh with a stubbed model loader."""
    import yaml
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree {REFERENCE_ROOT} not present (GPU box?) — use committed goldens")
    settings = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, "settings.yaml")))
    settings.update({
        "Depth Model": f"Depth-Anything-V2-{variant}",
        "Depth Resolution": depth_resolution,
        "FP16": fp16,
        "Display Mode": display_mode,
        "Run Mode": "Legacy Streamer",
        "torch.compile": None, "TensorRT": None, "CoreML": None, "OpenVINO": None, "MIGraphX": None,
        "Language": "EN",
    })
    if extra_settings:
        settings.update(extra_settings)
    work = tempfile.mkdtemp(prefix="d2s_ref_")
    with open(os.path.join(work, "settings.yaml"), "w") as f:
        yaml.safe_dump(settings, f)
    old_cwd = os.getcwd()
    os.chdir(work)
    try:
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        for m in ("depth", "utils"):
            sys.modules.pop(m, None)
        import transformers

        def _stub(*_a, **_k):
            return make_hf_model(variant, seed, tiny)

        transformers.AutoModelForDepthEstimation.from_pretrained = staticmethod(_stub)
        import depth  # noqa: E402  (the reference module)
    finally:
        os.chdir(old_cwd)
    return depth
