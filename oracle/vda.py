"""ORACLE (test infrastructure only) — fp32 functional restatement of streaming Video-Depth-Anything.

Follows the reference's vendored model (all under /root/reference/models/video_depth_anything/):
    ViT encoder          dinov2.py:179-210 (pos-embed: bicubic with scale_factor, +0.1 offset), :212-231, :275-318
                         (get_intermediate_layers: taps after blocks [2,5,8,11] / [4,11,17,23], final norm, cls dropped);
                         dinov2_layers/attention.py:44-62 (fused qkv, q pre-scaled, eager softmax); block.py (LayerScale ls1/ls2)
    DPT + temporal head  dpt_temporal.py:62-138, dpt.py:56-118, util/blocks.py:40-162
    temporal module      motion_module/motion_module.py:68-134 (GroupNorm(32, eps 1e-6) -> proj_in -> block -> proj_out -> +res),
                         :137-187 (2 x [LN -> temporal attention -> +res], LN -> GEGLU FF -> +res), :212-321 (attention over the
                         frame axis with sinusoidal APE; with a cache, Q is the newest frame only),
                         motion_module/attention.py:182-211 (_attention), :363-384 (GEGLU)
    streaming state      vda2_s.py:177-224 (first frame: cache = its hidden states x 31; later: shift left, append)
Pinned on the reference module itself run in fp32 on CPU (tests/golden/vda.npz via oracle/gen_golden.py)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from desktop2stereo_b200.synth import INFER_LEN, VDA_ENCODERS as ENCODERS, make_vda_state_dict as make_state_dict, param_shapes, sinusoid_pe  # noqa: E402,F401


def vit_features(sd, cfg, x):
    """x [T,3,H,W] -> 4 x [T, ph*pw, D] (normed, cls dropped)."""
    D, heads = cfg["hidden"], cfg["heads"]
    T, _, H, W = x.shape
    ph, pw = H // 14, W // 14
    t = F.conv2d(x, sd["pretrained.patch_embed.proj.weight"], sd["pretrained.patch_embed.proj.bias"], stride=14)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["pretrained.cls_token"].expand(T, -1, -1), t], 1)
    pos = sd["pretrained.pos_embed"].float()
    g = int(round((pos.shape[1] - 1) ** 0.5))
    if not (ph * pw == g * g and H == W):
        sx, sy = float(ph + 0.1) / g, float(pw + 0.1) / g
        pp = F.interpolate(pos[:, 1:].reshape(1, g, g, D).permute(0, 3, 1, 2), scale_factor=(sx, sy), mode="bicubic", antialias=False)
        assert pp.shape[-2:] == (ph, pw)
        pos = torch.cat([pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, D)], 1)
    t = t + pos
    feats = []
    for l in range(cfg["layers"]):
        p = f"pretrained.blocks.{l}."
        y = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(T, -1, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (D // heads) ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(-1) @ v
        a = F.linear(a.transpose(1, 2).reshape(T, -1, D), sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        t = t + a * sd[p + "ls1.gamma"]
        y = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        t = t + y * sd[p + "ls2.gamma"]
        if l in cfg["taps"]:
            feats.append(F.layer_norm(t, (D,), sd["pretrained.norm.weight"], sd["pretrained.norm.bias"], 1e-6)[:, 1:])
    return feats


def temporal_module(sd, m, x, caches, taps=None):
    """x [T,C,h,w] (T = 1 when streaming); caches: None or 2 tensors [(h*w), 31, C].  Returns (out, [2 new hidden states])."""
    t = f"head.motion_modules.{m}.temporal_transformer."
    T, C, h, w = x.shape
    hs = F.group_norm(x, 32, sd[t + "norm.weight"], sd[t + "norm.bias"], 1e-6)
    hs = hs.permute(0, 2, 3, 1).reshape(T, h * w, C)
    hs = F.linear(hs, sd[t + "proj_in.weight"], sd[t + "proj_in.bias"])
    b = t + "transformer_blocks.0."
    new = []
    heads = 8
    for a in range(2):
        ab = b + f"attention_blocks.{a}."
        n = F.layer_norm(hs, (C,), sd[b + f"norms.{a}.weight"], sd[b + f"norms.{a}.bias"])
        cur = n.permute(1, 0, 2)                                   # "(b f) d c -> (b d) f c"
        new.append(cur)
        d_in = 0
        seq = cur
        if caches is not None:
            d_in = caches[a].shape[1]
            seq = torch.cat([caches[a], cur], 1)
        seq = seq + sd[ab + "pos_encoder.pe"][:, :seq.shape[1]]
        q = F.linear(seq[:, d_in:], sd[ab + "to_q.weight"])
        k = F.linear(seq, sd[ab + "to_k.weight"])
        v = F.linear(seq, sd[ab + "to_v.weight"])
        sh = lambda z: z.reshape(z.shape[0], z.shape[1], heads, C // heads).permute(0, 2, 1, 3)
        att = (sh(q) @ sh(k).transpose(-1, -2) * (C // heads) ** -0.5).softmax(-1) @ sh(v)
        att = att.permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[1], C)
        att = F.linear(att, sd[ab + "to_out.0.weight"], sd[ab + "to_out.0.bias"])
        hs = att.permute(1, 0, 2) + hs                              # "(b d) f c -> (b f) d c"
    n = F.layer_norm(hs, (C,), sd[b + "ff_norm.weight"], sd[b + "ff_norm.bias"])
    val, gate = F.linear(n, sd[b + "ff.net.0.proj.weight"], sd[b + "ff.net.0.proj.bias"]).chunk(2, -1)
    hs = F.linear(val * F.gelu(gate), sd[b + "ff.net.2.weight"], sd[b + "ff.net.2.bias"]) + hs
    hs = F.linear(hs, sd[t + "proj_out.weight"], sd[t + "proj_out.bias"])
    out = hs.reshape(T, h, w, C).permute(0, 3, 1, 2) + x
    if taps is not None:
        taps[f"temporal{m}"] = out
    return out, new


def head_forward(sd, cfg, feats, ph, pw, caches, taps=None):
    """dpt_temporal.py:62-138 for one micro-batch; caches: None or 8 tensors.  Returns depth [T,1,H,W] and 8 new hidden states."""
    T = feats[0].shape[0]
    D = cfg["hidden"]
    maps = []
    for i, f in enumerate(feats):
        m = f.permute(0, 2, 1).reshape(T, D, ph, pw)
        m = F.conv2d(m, sd[f"head.projects.{i}.weight"], sd[f"head.projects.{i}.bias"])
        if i == 0:
            m = F.conv_transpose2d(m, sd["head.resize_layers.0.weight"], sd["head.resize_layers.0.bias"], stride=4)
        elif i == 1:
            m = F.conv_transpose2d(m, sd["head.resize_layers.1.weight"], sd["head.resize_layers.1.bias"], stride=2)
        elif i == 3:
            m = F.conv2d(m, sd["head.resize_layers.3.weight"], sd["head.resize_layers.3.bias"], stride=2, padding=1)
        maps.append(m)
    c = (lambda i: caches[2 * i:2 * i + 2]) if caches is not None else (lambda i: None)
    maps[2], h0 = temporal_module(sd, 0, maps[2], c(0), taps)
    maps[3], h1 = temporal_module(sd, 1, maps[3], c(1), taps)
    rn = [F.conv2d(m, sd[f"head.scratch.layer{i + 1}_rn.weight"], None, padding=1) for i, m in enumerate(maps)]

    def rcu(x, pre):
        y = F.conv2d(F.relu(x), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
        y = F.conv2d(F.relu(y), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
        return y + x

    def fusion(r, x0, x1, size):
        p = f"head.scratch.refinenet{r}."
        o = x0 if x1 is None else x0 + rcu(x1, p + "resConfUnit1.")
        o = rcu(o, p + "resConfUnit2.")
        o = F.interpolate(o, size=size, mode="bilinear", align_corners=True) if size is not None else \
            F.interpolate(o, scale_factor=2, mode="bilinear", align_corners=True)
        return F.conv2d(o, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])

    p4 = fusion(4, rn[3], None, rn[2].shape[2:])
    p4, h2 = temporal_module(sd, 2, p4, c(2), taps)
    p3 = fusion(3, p4, rn[2], rn[1].shape[2:])
    p3, h3 = temporal_module(sd, 3, p3, c(3), taps)
    p2 = fusion(2, p3, rn[1], rn[0].shape[2:])
    p1 = fusion(1, p2, rn[0], None)
    o = F.conv2d(p1, sd["head.scratch.output_conv1.weight"], sd["head.scratch.output_conv1.bias"], padding=1)
    o = F.interpolate(o, (ph * 14, pw * 14), mode="bilinear", align_corners=True)
    o = F.relu(F.conv2d(o, sd["head.scratch.output_conv2.0.weight"], sd["head.scratch.output_conv2.0.bias"], padding=1))
    o = F.relu(F.conv2d(o, sd["head.scratch.output_conv2.2.weight"], sd["head.scratch.output_conv2.2.bias"]))
    return F.relu(o), h0 + h1 + h2 + h3   # vda2_s.py:83-84 (the resize to (H,W) is the identity: H == 14*ph)


class StreamingVDA:
    """vda2_s.py:189-224: one frame per call; state = 8 caches of the last 31 frames' normed hidden states."""

    def __init__(self, sd: dict, encoder: str):
        self.sd, self.cfg = sd, ENCODERS[encoder]
        self.cache = None

    def reset(self):
        self.cache = None

    @torch.no_grad()
    def __call__(self, pixel_values: torch.Tensor, taps=None) -> torch.Tensor:
        """pixel_values [1,3,H,W] fp32 -> depth [1,1,H,W]"""
        x = pixel_values.float()
        _, _, H, W = x.shape
        feats = vit_features(self.sd, self.cfg, x)
        if taps is not None:
            for i, f in enumerate(feats):
                taps[f"feat{i}"] = f
        depth, new = head_forward(self.sd, self.cfg, feats, H // 14, W // 14, self.cache, taps)
        if self.cache is None:
            self.cache = [torch.cat([h] * (INFER_LEN - 1), 1).contiguous() for h in new]
        else:
            self.cache = [torch.cat([c[:, 1:], h], 1) for c, h in zip(self.cache, new)]
        return depth
