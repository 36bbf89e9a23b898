"""ORACLE (test infrastructure) — fp32 functional restatement of Depth Anything V2's forward pass.

The arithmetic lives in a third-party dependency of the reference, HF `transformers` (pinned 4.56.2 in the
reference's requirements.txt:5; 5.5.0 installed here), not under /root/reference.  Call sites in the reference:
depth.py:14 (import), :1649-1662 (from_pretrained), :1778 (`self.model(pixel_values=tensor).predicted_depth`).
This file restates the published algorithm from a plain state_dict:
    embeddings        HF models/dinov2/modeling_dinov2.py:38-149
    encoder layer     :182-386   (LN -> q,k,v -> softmax(QK^T/8)V -> dense -> LayerScale -> +res; LN -> fc1 -> GELU(erf) -> fc2 -> LayerScale -> +res)
    backbone taps     :605-618   (shared final LayerNorm on hidden_states[out_indices], cls kept)
    reassemble        HF models/depth_anything/modeling_depth_anything.py:31-93
    fusion            :96-203
    head              :292-308
It is pinned on outputs of the HF model itself (tests/golden/model.npz, tests/test_oracle_model.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def forward(sd: dict, cfg: dict, pixel_values: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """sd: HF state_dict (fp32 tensors).  cfg: dict(hidden, layers, heads, out_indices, neck, fusion, patch, eps,
    metric, max_depth).  pixel_values [B,3,H,W] fp32.  Returns predicted_depth [B,H,W]."""
    D, L, heads, patch = cfg["hidden"], cfg["layers"], cfg["heads"], cfg.get("patch", 14)
    eps = cfg.get("eps", 1e-6)
    x = pixel_values.float()
    B, _, H, W = x.shape
    ph, pw = H // patch, W // patch
    e = "backbone.embeddings."
    t = F.conv2d(x, sd[e + "patch_embeddings.projection.weight"], sd[e + "patch_embeddings.projection.bias"], stride=patch)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd[e + "cls_token"].expand(B, -1, -1), t], 1)
    pos = sd[e + "position_embeddings"]
    g = int(round((pos.shape[1] - 1) ** 0.5))
    if not (ph == g and pw == g):
        pp = pos[:, 1:].reshape(1, g, g, D).permute(0, 3, 1, 2)
        pp = F.interpolate(pp.float(), size=(ph, pw), mode="bicubic", align_corners=False)
        pos = torch.cat([pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, D)], 1)
    t = t + pos
    if taps is not None:
        taps["embeddings"] = t
    feats = []
    for l in range(L):
        p = f"backbone.encoder.layer.{l}."
        a = p + "attention.attention."
        y = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        q = F.linear(y, sd[a + "query.weight"], sd[a + "query.bias"]).view(B, -1, heads, D // heads).transpose(1, 2)
        k = F.linear(y, sd[a + "key.weight"], sd[a + "key.bias"]).view(B, -1, heads, D // heads).transpose(1, 2)
        v = F.linear(y, sd[a + "value.weight"], sd[a + "value.bias"]).view(B, -1, heads, D // heads).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2) * (D // heads) ** -0.5, dim=-1) @ v
        att = att.transpose(1, 2).reshape(B, -1, D)
        att = F.linear(att, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        t = att * sd[p + "layer_scale1.lambda1"] + t
        y = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        t = y * sd[p + "layer_scale2.lambda1"] + t
        if (l + 1) in cfg["out_indices"]:
            feats.append(F.layer_norm(t, (D,), sd["backbone.layernorm.weight"], sd["backbone.layernorm.bias"], eps))
    if taps is not None:
        taps["hidden_last"] = t
        for i, f in enumerate(feats):
            taps[f"feat{i}"] = f[:, 1:]
    # reassemble
    r = "neck.reassemble_stage.layers."
    maps = []
    for i, f in enumerate(feats):
        m = f[:, 1:].reshape(B, ph, pw, D).permute(0, 3, 1, 2)
        m = F.conv2d(m, sd[f"{r}{i}.projection.weight"], sd[f"{r}{i}.projection.bias"])
        if i == 0:
            m = F.conv_transpose2d(m, sd[f"{r}0.resize.weight"], sd[f"{r}0.resize.bias"], stride=4)
        elif i == 1:
            m = F.conv_transpose2d(m, sd[f"{r}1.resize.weight"], sd[f"{r}1.resize.bias"], stride=2)
        elif i == 3:
            m = F.conv2d(m, sd[f"{r}3.resize.weight"], sd[f"{r}3.resize.bias"], stride=2, padding=1)
        maps.append(m)
        if taps is not None:
            taps[f"reassemble{i}"] = m
    maps = [F.conv2d(m, sd[f"neck.convs.{i}.weight"], None, padding=1) for i, m in enumerate(maps)]
    if taps is not None:
        for i, m in enumerate(maps):
            taps[f"neck{i}"] = m

    def rcu(x, pre):
        y = F.conv2d(F.relu(x), sd[pre + "convolution1.weight"], sd[pre + "convolution1.bias"], padding=1)
        y = F.conv2d(F.relu(y), sd[pre + "convolution2.weight"], sd[pre + "convolution2.bias"], padding=1)
        return y + x

    fused = None
    rev = maps[::-1]
    for j, m in enumerate(rev):
        f = f"neck.fusion_stage.layers.{j}."
        hs = m if fused is None else fused + rcu(m, f + "residual_layer1.")
        hs = rcu(hs, f + "residual_layer2.")
        if j != len(rev) - 1:
            hs = F.interpolate(hs, size=rev[j + 1].shape[2:], mode="bilinear", align_corners=True)
        else:
            hs = F.interpolate(hs, scale_factor=2, mode="bilinear", align_corners=True)
        fused = F.conv2d(hs, sd[f + "projection.weight"], sd[f + "projection.bias"])
        if taps is not None:
            taps[f"fused{j}"] = fused
    d = F.conv2d(fused, sd["head.conv1.weight"], sd["head.conv1.bias"], padding=1)
    if taps is not None:
        taps["head_conv1"] = d
    d = F.interpolate(d, (ph * patch, pw * patch), mode="bilinear", align_corners=True)
    d = F.relu(F.conv2d(d, sd["head.conv2.weight"], sd["head.conv2.bias"], padding=1))
    d = F.conv2d(d, sd["head.conv3.weight"], sd["head.conv3.bias"])
    d = torch.sigmoid(d) * cfg.get("max_depth", 1.0) if cfg.get("metric") else F.relu(d)
    return d.squeeze(1)


def cfg_from_hf(hf_config) -> dict:
    b = hf_config.backbone_config
    return dict(hidden=b.hidden_size, layers=b.num_hidden_layers, heads=b.num_attention_heads,
                out_indices=list(b.out_indices), neck=list(hf_config.neck_hidden_sizes),
                fusion=hf_config.fusion_hidden_size, patch=b.patch_size, eps=b.layer_norm_eps,
                metric=hf_config.depth_estimation_type == "metric", max_depth=hf_config.max_depth or 1.0)
