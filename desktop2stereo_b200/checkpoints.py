"""Checkpoint files -> packed weight blob, without instantiating a PyTorch module.

    load_hf_checkpoint(dir)            a Hugging Face `-hf` Depth-Anything-V2 snapshot: config.json + model.safetensors
                                       (what AutoModelForDepthEstimation.from_pretrained reads, reference depth.py:1646-1662)
    load_vda_checkpoint(path, encoder) video_depth_anything_{vits,vitb,vitl}.pth (torch.load of a plain state_dict,
                                       reference depth.py:885-901)
Both return (blob, ModelConfig) ready for B200Engine(blob, cfg, device).  Host-side only.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .weights import config_for_vda, config_from_hf, pack_state_dict, pack_vda_state_dict


def load_hf_checkpoint(path: str):
    from safetensors import safe_open
    from transformers import DepthAnythingConfig
    with open(os.path.join(path, "config.json")) as f:
        hf_config = DepthAnythingConfig.from_dict(json.load(f))
    cfg = config_from_hf(hf_config)
    sd = {}
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors file under {path}")
    for name in files:
        # framework="pt": fp16 AND bf16 snapshots load (numpy has no bfloat16); everything is widened to fp32 for packing
        with safe_open(os.path.join(path, name), framework="pt") as st:
            for k in st.keys():
                sd[k] = st.get_tensor(k).float().numpy()
    return pack_state_dict(sd, cfg), cfg


def load_vda_checkpoint(path: str, encoder: str):
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=True)
    cfg = config_for_vda(encoder)
    return pack_vda_state_dict(sd, cfg), cfg
