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

Reference loading path this stands in for: depth.py:1633-1690 (_load_pytorch_model) + HF from_pretrained.
"""
from __future__ import annotations

import numpy as np

from ._lib import ModelConfig


def config_from_hf(hf_config, max_batch=1, max_h=518, max_w=518) -> ModelConfig:
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
