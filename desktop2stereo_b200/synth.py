"""Seeded synthetic weights for the benchmarks and tests (there are no checkpoints offline).

    make_hf_model(variant, seed, tiny)      a random-init transformers DepthAnythingForDepthEstimation (DA-V2 S/B/L or a tiny config)
    make_vda_state_dict(encoder, seed)      a state_dict with the names/shapes of the reference's VideoDepthAnything (vda2_s.py)
Parameters that the stock initialisers set to constants (LayerScale = 1, biases = 0, zero-initialised proj_out) are re-drawn so that
every term of the forward pass is exercised.  Host-side only; nothing here touches the GPU path.
"""
from __future__ import annotations

import math

import torch

# name -> (hidden, layers, heads, out_indices, neck_hidden_sizes, fusion_hidden_size)
# depth.py:889-893 (VDA table) and the HF `-hf` checkpoints' config.json.
DA_V2_VARIANTS = {
    "Small": (384, 12, 6, [3, 6, 9, 12], [48, 96, 192, 384], 64),
    "Base": (768, 12, 12, [3, 6, 9, 12], [96, 192, 384, 768], 128),
    "Large": (1024, 24, 16, [5, 12, 18, 24], [256, 512, 1024, 1024], 256),
}


# a small DA-V2 configuration for fast parity tests (the architecture, scaled down)
TINY_CFG = dict(hidden=128, layers=4, heads=2, out_indices=[1, 2, 3, 4], neck=[24, 48, 96, 192], fusion=64)


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



INFER_LEN = 32
VDA_ENCODERS = {  # depth.py:889-893 + dinov2.py:339-377 + vda2_s.py:52-56
    "vits": dict(hidden=384, layers=12, heads=6, taps=[2, 5, 8, 11], features=64, out_channels=[48, 96, 192, 384]),
    "vitb": dict(hidden=768, layers=12, heads=12, taps=[2, 5, 8, 11], features=128, out_channels=[96, 192, 384, 768]),
    "vitl": dict(hidden=1024, layers=24, heads=16, taps=[4, 11, 17, 23], features=256, out_channels=[256, 512, 1024, 1024]),
}


def param_shapes(encoder: str):
    """(name, shape) of every tensor in VideoDepthAnything(encoder).state_dict(), in module order."""
    c = VDA_ENCODERS[encoder]
    D, F_, oc = c["hidden"], c["features"], c["out_channels"]
    s = [("pretrained.cls_token", (1, 1, D)), ("pretrained.pos_embed", (1, 1370, D)), ("pretrained.mask_token", (1, D)),
         ("pretrained.patch_embed.proj.weight", (D, 3, 14, 14)), ("pretrained.patch_embed.proj.bias", (D,))]
    for l in range(c["layers"]):
        p = f"pretrained.blocks.{l}."
        s += [(p + "norm1.weight", (D,)), (p + "norm1.bias", (D,)), (p + "attn.qkv.weight", (3 * D, D)), (p + "attn.qkv.bias", (3 * D,)),
              (p + "attn.proj.weight", (D, D)), (p + "attn.proj.bias", (D,)), (p + "ls1.gamma", (D,)),
              (p + "norm2.weight", (D,)), (p + "norm2.bias", (D,)), (p + "mlp.fc1.weight", (4 * D, D)), (p + "mlp.fc1.bias", (4 * D,)),
              (p + "mlp.fc2.weight", (D, 4 * D)), (p + "mlp.fc2.bias", (D,)), (p + "ls2.gamma", (D,))]
    s += [("pretrained.norm.weight", (D,)), ("pretrained.norm.bias", (D,))]
    for i in range(4):
        s += [(f"head.projects.{i}.weight", (oc[i], D, 1, 1)), (f"head.projects.{i}.bias", (oc[i],))]
    s += [("head.resize_layers.0.weight", (oc[0], oc[0], 4, 4)), ("head.resize_layers.0.bias", (oc[0],)),
          ("head.resize_layers.1.weight", (oc[1], oc[1], 2, 2)), ("head.resize_layers.1.bias", (oc[1],)),
          ("head.resize_layers.3.weight", (oc[3], oc[3], 3, 3)), ("head.resize_layers.3.bias", (oc[3],))]
    for i in range(4):
        s.append((f"head.scratch.layer{i + 1}_rn.weight", (F_, oc[i], 3, 3)))
    for r in (1, 2, 3, 4):
        p = f"head.scratch.refinenet{r}."
        s += [(p + "out_conv.weight", (F_, F_, 1, 1)), (p + "out_conv.bias", (F_,))]
        for u in ("resConfUnit1.", "resConfUnit2."):
            for cv in ("conv1.", "conv2."):
                s += [(p + u + cv + "weight", (F_, F_, 3, 3)), (p + u + cv + "bias", (F_,))]
    s += [("head.scratch.output_conv1.weight", (F_ // 2, F_, 3, 3)), ("head.scratch.output_conv1.bias", (F_ // 2,)),
          ("head.scratch.output_conv2.0.weight", (32, F_ // 2, 3, 3)), ("head.scratch.output_conv2.0.bias", (32,)),
          ("head.scratch.output_conv2.2.weight", (1, 32, 1, 1)), ("head.scratch.output_conv2.2.bias", (1,))]
    for m, C in enumerate([oc[2], oc[3], F_, F_]):
        t = f"head.motion_modules.{m}.temporal_transformer."
        s += [(t + "norm.weight", (C,)), (t + "norm.bias", (C,)), (t + "proj_in.weight", (C, C)), (t + "proj_in.bias", (C,))]
        b = t + "transformer_blocks.0."
        for a in range(2):
            ab = b + f"attention_blocks.{a}."
            s += [(ab + "to_q.weight", (C, C)), (ab + "to_k.weight", (C, C)), (ab + "to_v.weight", (C, C)),
                  (ab + "to_out.0.weight", (C, C)), (ab + "to_out.0.bias", (C,)), (ab + "pos_encoder.pe", (1, INFER_LEN, C))]
        for a in range(2):
            s += [(b + f"norms.{a}.weight", (C,)), (b + f"norms.{a}.bias", (C,))]
        s += [(b + "ff.net.0.proj.weight", (8 * C, C)), (b + "ff.net.0.proj.bias", (8 * C,)),
              (b + "ff.net.2.weight", (C, 4 * C)), (b + "ff.net.2.bias", (C,)),
              (b + "ff_norm.weight", (C,)), (b + "ff_norm.bias", (C,)),
              (t + "proj_out.weight", (C, C)), (t + "proj_out.bias", (C,))]
    return s


def sinusoid_pe(C: int, max_len: int = INFER_LEN) -> torch.Tensor:
    """motion_module.py:190-204"""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, C, 2) * (-math.log(10000.0) / C))
    pe = torch.zeros(1, max_len, C)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def make_vda_state_dict(encoder: str, seed: int) -> dict:
    """Seeded, variance-preserving random weights (regenerable on any box without the reference): every term of the forward
    pass is exercised (the shipped init has proj_out == 0, LayerScale == 1, biases == 0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(encoder):
        if name.endswith("pos_encoder.pe"):
            sd[name] = sinusoid_pe(shape[2])
        elif name.endswith(".gamma"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("cls_token") or name.endswith("pos_embed") or name.endswith("mask_token"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif name == "head.scratch.output_conv2.2.bias":
            sd[name] = torch.full(shape, 0.5)
        elif name.endswith("weight") and len(shape) >= 2:
            if "resize_layers.0" in name or "resize_layers.1" in name:
                fan_in = shape[0]                      # ConvTranspose with kernel == stride: one tap per output pixel
            else:
                fan_in = int(torch.tensor(shape[1:]).prod())
            sd[name] = torch.randn(shape, generator=g) / fan_in ** 0.5
        elif name.endswith("weight"):                  # norm scales
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                          # biases
            sd[name] = 0.05 * torch.randn(shape, generator=g)
    return sd


