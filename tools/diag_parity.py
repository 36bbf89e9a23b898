#!/usr/bin/env python
"""GPU-box diagnostic: what does the REFERENCE's own CUDA numerics score against the fp32 oracle?

The reference runs the network under `torch.autocast("cuda")` => fp16 GEMMs/convs/SDPA with fp32 LayerNorm/softmax
(reference depth.py:1763-1781, 661-664).  This script runs, on the same seeded weights and inputs,
    (a) HF DepthAnythingForDepthEstimation in fp32 (TF32 off)                       -> the oracle value
    (b) the same module under torch.autocast("cuda", dtype=float16)                 -> the reference's CUDA path
    (c) the B200 engine (d2s_infer through the C ABI), latency and throughput plans
and prints max|d - d32| / max|d32| and mean|d - d32| / max|d32| for (b) and (c) side by side, at the BASELINE configs'
own network shapes.  VDA: the fp32 functional restatement (oracle/vda.py) vs the same restatement under fp16 autocast vs
the temporal engine, streamed for 34 frames.  Writes JSON lines to stdout and gpurun_out/parity_<tag>.jsonl.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from desktop2stereo_b200.engine import B200Engine          # noqa: E402
from desktop2stereo_b200.synth import make_hf_model          # noqa: E402
from oracle import vda                                         # noqa: E402
from oracle.gen_golden import model_input, vda_frames          # noqa: E402


def errs(a, ref):
    d = (a.float() - ref).abs()
    m = ref.abs().max().item()
    return {"max": d.max().item() / m, "mean": d.mean().item() / m}


def dav2_case(name, variant, seed, B, H, W, dev, out):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = make_hf_model(variant, seed)
    x = torch.from_numpy(model_input(seed, B, H, W)).to(dev)
    eng = B200Engine.from_hf_model(model, dev, out_dtype=torch.float32)
    model = model.to(dev)
    with torch.no_grad():
        ref = model(pixel_values=x).predicted_depth
        with torch.autocast("cuda", dtype=torch.float16):
            r16 = model(pixel_values=x).predicted_depth
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rb16 = model(pixel_values=x).predicted_depth
    lat = eng(x).clone()
    eng.set_policy("throughput")
    thr = eng(x).clone()
    rec = {"case": name, "variant": variant, "B": B, "H": H, "W": W, "ref_max": ref.abs().max().item(),
           "frac_pos": (ref > 0).float().mean().item(), "ref_dtype_under_autocast": str(r16.dtype),
           "reference_fp16_autocast": errs(r16, ref), "reference_bf16_autocast": errs(rb16, ref),
           "engine_latency": errs(lat, ref), "engine_throughput": errs(thr, ref)}
    eng.close()
    del model
    torch.cuda.empty_cache()
    print(json.dumps(rec), flush=True)
    out.append(rec)


def vda_case(encoder, seed, T, H, W, dev, out):
    sd = vda.make_state_dict(encoder, seed)
    sdd = {k: v.to(dev) for k, v in sd.items()}
    eng = B200Engine.from_vda_state_dict(sd, encoder, dev, out_dtype=torch.float32)
    o32 = vda.StreamingVDA(sdd, encoder)
    o16 = vda.StreamingVDA(sdd, encoder)
    frames = torch.from_numpy(vda_frames(seed, T, H, W)).to(dev)
    w16 = {"max": 0.0, "mean": 0.0}
    we = {"max": 0.0, "mean": 0.0}
    for t in range(T):
        with torch.no_grad():
            ref = o32(frames[t])
            with torch.autocast("cuda", dtype=torch.float16):
                r16 = o16(frames[t])
        got = eng(frames[t])
        a, b = errs(r16, ref), errs(got, ref)
        for k in ("max", "mean"):
            w16[k] = max(w16[k], a[k]); we[k] = max(we[k], b[k])
    rec = {"case": f"vda_{encoder}_{H}x{W}", "frames": T, "ref_max": ref.abs().max().item(), "frac_pos": (ref > 0).float().mean().item(),
           "reference_fp16_autocast(worst frame)": w16, "engine(worst frame)": we}
    eng.close()
    torch.cuda.empty_cache()
    print(json.dumps(rec), flush=True)
    out.append(rec)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rX"
    what = sys.argv[2:] or ["small518", "base1", "base2", "large8", "vits", "vitl"]
    dev = torch.device("cuda:0")
    out = []
    if "small518" in what:
        dav2_case("small_518x518_b1 (config 1)", "Small", 5, 1, 518, 518, dev, out)
    if "base1" in what:
        dav2_case("base_294x518_b1 (config 2)", "Base", 7, 1, 294, 518, dev, out)
    if "base2" in what:
        dav2_case("base_294x518_b2", "Base", 7, 2, 294, 518, dev, out)
    if "large8" in what:
        dav2_case("large_294x518_b8 (configs 3/5)", "Large", 9, 8, 294, 518, dev, out)
    if "large1" in what:
        dav2_case("large_294x518_b1", "Large", 9, 1, 294, 518, dev, out)
    if "vits" in what:
        vda_case("vits", 21, 34, 294, 518, dev, out)
    if "vitl" in what:
        vda_case("vitl", 22, 34, 294, 518, dev, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_{tag}.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
