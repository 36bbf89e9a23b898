"""Pack a Depth-Anything-V2 (HF `DepthAnythingForDepthEstimation`) state_dict into the flat fp32 blob
`d2s_create` consumes.  Host-side numpy only.

Blob order (all fp32, row-major, nn.Module layouts unless noted):
    patch_w [D,588] patch_b [D] cls [D] pos [1+g*g, D]
    per encoder layer: ln1_w ln1_b | qkv_w [3D,D] (q;k;v) qkv_b [3D] | proj_w [D,D]* proj_b [D]* |
                       ln2_w ln2_b | fc1_w [4D,D] fc1_b | fc2_w [D,4D]* fc2_b [D]*        (* LayerScale folded in)
    norm_w norm_b                                                  (backbone.layernorm, applied to the 4 taps)
    reassemble: 4 x (proj_w [c_i,D], proj_b) | up0 ConvT w [c0,c0,4,4] b | up1 ConvT w [c1,c1,2,2] b | down3 Conv w [c3,c3,3,3] b
    neck convs: 4 x w [F,c_i,3,3]
    fusion layers j=0..3: proj_w [F,F] proj_b | rl1.conv1 w b | rl1.conv2 w b | rl2.conv1 w b | rl2.conv2 w b
    head: conv1 w [F/2,F,3,3] b | conv2 w [32,F/2,3,3] b | conv3 w [32] b [1]

    temporal engines (Video-Depth-Anything) append, for each of the 4 TemporalModules (C = c2, c3, F, F):
        gn_w gn_b | proj_in_w [C,C] b | 2 x (ln_w ln_b | qkv_w [3C,C] (to_q;to_k;to_v, bias-free) | pe_qkv [32,3C] = pe @ qkv_w^T |
        out_w [C,C] out_b) | ffln_w ffln_b | ff1_w [8C,C] b | ff2_w [C,4C] b | proj_out_w [C,C] b

Reference loading path this stands in for: depth.py:1633-1690 (_load_pytorch_model) + HF from_pretrained, and
depth.py:870-902 (get_video_depth_anything_model: torch.load of video_depth_anything_{vits,vitb,vitl}.pth).
"""
from __future__ import annotations

import numpy as np

from ._lib import ModelConfig


def config_from_hf(hf_config, max_batch=0, max_h=0, max_w=0) -> ModelConfig:
    b = hf_config.backbone_config
    c = ModelConfig()
    c.hidden, c.layers, c.heads = b.hidden_size, b.num_hidden_layers, b.num_attention_heads
    c.mlp_hidden = int(b.hidden_size * b.mlp_ratio)
    c.patch = b.patch_size
    c.pos_grid = b.image_size // b.patch_size
    idx = list(getattr(b, "out_indices", None) or [])
    if len(idx) != 4:
        raise ValueError(f"need 4 backbone taps, got {idx}")
    if getattr(b, "use_swiglu_ffn", False):
        raise ValueError("SwiGLU FFN (DINOv2-giant) is not supported")
    for i in range(4):
        c.out_indices[i] = int(idx[i])
        c.neck[i] = int(hf_config.neck_hidden_sizes[i])
    if list(hf_config.reassemble_factors) != [4, 2, 1, 0.5]:
        raise ValueError(f"unsupported reassemble_factors {hf_config.reassemble_factors}")
    c.fusion, c.head_hidden = hf_config.fusion_hidden_size, hf_config.head_hidden_size
    c.layer_norm_eps = b.layer_norm_eps
    c.metric = int(hf_config.depth_estimation_type == "metric")
    c.max_depth = float(hf_config.max_depth if hf_config.max_depth is not None else 1.0)
    c.max_batch, c.max_h, c.max_w = max_batch, max_h, max_w
    return c


def pack_state_dict(sd: dict, cfg: ModelConfig) -> np.ndarray:
    """sd: name -> array-like (torch tensors are accepted); returns the fp32 blob."""

    def g(name):
        t = sd[name]
        if hasattr(t, "detach"):
            t = t.detach().to("cpu").float().numpy()
        return np.ascontiguousarray(t, dtype=np.float32)

    D, L = cfg.hidden, cfg.layers
    out = []
    emb = "backbone.embeddings."
    out += [g(emb + "patch_embeddings.projection.weight").reshape(D, -1), g(emb + "patch_embeddings.projection.bias"),
            g(emb + "cls_token").reshape(D), g(emb + "position_embeddings").reshape(-1, D)]
    for l in range(L):
        p = f"backbone.encoder.layer.{l}."
        ls1, ls2 = g(p + "layer_scale1.lambda1"), g(p + "layer_scale2.lambda1")
        a = p + "attention.attention."
        qkv_w = np.concatenate([g(a + "query.weight"), g(a + "key.weight"), g(a + "value.weight")], 0)
        qkv_b = np.concatenate([g(a + "query.bias"), g(a + "key.bias"), g(a + "value.bias")], 0)
        out += [g(p + "norm1.weight"), g(p + "norm1.bias"), qkv_w, qkv_b,
                g(p + "attention.output.dense.weight") * ls1[:, None], g(p + "attention.output.dense.bias") * ls1,
                g(p + "norm2.weight"), g(p + "norm2.bias"),
                g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias"),
                g(p + "mlp.fc2.weight") * ls2[:, None], g(p + "mlp.fc2.bias") * ls2]
    out += [g("backbone.layernorm.weight"), g("backbone.layernorm.bias")]
    r = "neck.reassemble_stage.layers."
    for i in range(4):
        out += [g(f"{r}{i}.projection.weight").reshape(cfg.neck[i], D), g(f"{r}{i}.projection.bias")]
    out += [g(f"{r}0.resize.weight"), g(f"{r}0.resize.bias"), g(f"{r}1.resize.weight"), g(f"{r}1.resize.bias"),
            g(f"{r}3.resize.weight"), g(f"{r}3.resize.bias")]
    for i in range(4):
        out.append(g(f"neck.convs.{i}.weight"))
    F = cfg.fusion
    for j in range(4):
        f = f"neck.fusion_stage.layers.{j}."
        out += [g(f + "projection.weight").reshape(F, F), g(f + "projection.bias")]
        for rl in ("residual_layer1.", "residual_layer2."):
            out += [g(f + rl + "convolution1.weight"), g(f + rl + "convolution1.bias"),
                    g(f + rl + "convolution2.weight"), g(f + rl + "convolution2.bias")]
    out += [g("head.conv1.weight"), g("head.conv1.bias"), g("head.conv2.weight"), g("head.conv2.bias"),
            g("head.conv3.weight").reshape(-1), g("head.conv3.bias").reshape(1)]
    return np.concatenate([a.reshape(-1) for a in out]).astype(np.float32, copy=False)


# Video-Depth-Anything encoders (depth.py:889-893; dinov2.py:339-377; vda2_s.py:52-56)
VDA_ENCODERS = {
    "vits": dict(hidden=384, layers=12, heads=6, taps=[2, 5, 8, 11], features=64, out_channels=[48, 96, 192, 384]),
    "vitb": dict(hidden=768, layers=12, heads=12, taps=[2, 5, 8, 11], features=128, out_channels=[96, 192, 384, 768]),
    "vitl": dict(hidden=1024, layers=24, heads=16, taps=[4, 11, 17, 23], features=256, out_channels=[256, 512, 1024, 1024]),
}


def config_for_vda(encoder: str, max_h=0, max_w=0) -> ModelConfig:
    e = VDA_ENCODERS[encoder]
    c = ModelConfig()
    c.hidden, c.layers, c.heads, c.mlp_hidden = e["hidden"], e["layers"], e["heads"], 4 * e["hidden"]
    c.patch, c.pos_grid = 14, 37
    for i in range(4):
        c.out_indices[i] = e["taps"][i] + 1          # block index (0-based) -> hidden-state index (1-based)
        c.neck[i] = e["out_channels"][i]
    c.fusion, c.head_hidden = e["features"], 32
    c.layer_norm_eps, c.metric, c.max_depth = 1e-6, 0, 1.0
    c.max_batch, c.max_h, c.max_w = 1, max_h, max_w
    c.temporal, c.pos_interp_offset = 1, 0.1
    return c


def pack_vda_state_dict(sd: dict, cfg: ModelConfig) -> np.ndarray:
    """sd: state_dict of the reference's VideoDepthAnything (vda2_s.py) -> the fp32 blob of a temporal engine."""

    def g(name):
        t = sd[name]
        if hasattr(t, "detach"):
            t = t.detach().to("cpu").float().numpy()
        return np.ascontiguousarray(t, dtype=np.float32)

    D, L, F = cfg.hidden, cfg.layers, cfg.fusion
    out = [g("pretrained.patch_embed.proj.weight").reshape(D, -1), g("pretrained.patch_embed.proj.bias"),
           g("pretrained.cls_token").reshape(D), g("pretrained.pos_embed").reshape(-1, D)]
    for l in range(L):
        p = f"pretrained.blocks.{l}."
        ls1, ls2 = g(p + "ls1.gamma"), g(p + "ls2.gamma")
        out += [g(p + "norm1.weight"), g(p + "norm1.bias"), g(p + "attn.qkv.weight"), g(p + "attn.qkv.bias"),
                g(p + "attn.proj.weight") * ls1[:, None], g(p + "attn.proj.bias") * ls1,
                g(p + "norm2.weight"), g(p + "norm2.bias"), g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias"),
                g(p + "mlp.fc2.weight") * ls2[:, None], g(p + "mlp.fc2.bias") * ls2]
    out += [g("pretrained.norm.weight"), g("pretrained.norm.bias")]
    for i in range(4):
        out += [g(f"head.projects.{i}.weight").reshape(cfg.neck[i], D), g(f"head.projects.{i}.bias")]
    for i in (0, 1, 3):
        out += [g(f"head.resize_layers.{i}.weight"), g(f"head.resize_layers.{i}.bias")]
    for i in range(4):
        out.append(g(f"head.scratch.layer{i + 1}_rn.weight"))
    for j in range(4):                                   # fusion order coarse -> fine: refinenet4 .. refinenet1
        f = f"head.scratch.refinenet{4 - j}."
        out += [g(f + "out_conv.weight").reshape(F, F), g(f + "out_conv.bias")]
        for u in ("resConfUnit1.", "resConfUnit2."):
            out += [g(f + u + "conv1.weight"), g(f + u + "conv1.bias"), g(f + u + "conv2.weight"), g(f + u + "conv2.bias")]
    out += [g("head.scratch.output_conv1.weight"), g("head.scratch.output_conv1.bias"),
            g("head.scratch.output_conv2.0.weight"), g("head.scratch.output_conv2.0.bias"),
            g("head.scratch.output_conv2.2.weight").reshape(-1), g("head.scratch.output_conv2.2.bias").reshape(1)]
    for m in range(4):
        t = f"head.motion_modules.{m}.temporal_transformer."
        b = t + "transformer_blocks.0."
        out += [g(t + "norm.weight"), g(t + "norm.bias"), g(t + "proj_in.weight"), g(t + "proj_in.bias")]
        for a in range(2):
            ab = b + f"attention_blocks.{a}."
            qkv = np.concatenate([g(ab + "to_q.weight"), g(ab + "to_k.weight"), g(ab + "to_v.weight")], 0)
            pe = g(ab + "pos_encoder.pe")[0].astype(np.float64)             # [32, C] sinusoid (motion_module.py:190-204)
            out += [g(b + f"norms.{a}.weight"), g(b + f"norms.{a}.bias"), qkv,
                    (pe @ qkv.astype(np.float64).T).astype(np.float32),      # position terms of q/k/v: the projections are linear
                    g(ab + "to_out.0.weight"), g(ab + "to_out.0.bias")]
        out += [g(b + "ff_norm.weight"), g(b + "ff_norm.bias"), g(b + "ff.net.0.proj.weight"), g(b + "ff.net.0.proj.bias"),
                g(b + "ff.net.2.weight"), g(b + "ff.net.2.bias"), g(t + "proj_out.weight"), g(t + "proj_out.bias")]
    return np.concatenate([a.reshape(-1) for a in out]).astype(np.float32, copy=False)
