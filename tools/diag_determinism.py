"""Replay the same input several times and report, tap by tap, the first activation that differs between replays."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from desktop2stereo_b200.engine import B200Engine
from oracle.gen_golden import MODEL_CASES, model_input
from oracle.ref_harness import make_hf_model

dev = torch.device("cuda:0")
TAPS = ["hidden_last"] + [f"feat{i}" for i in range(4)] + [f"reassemble{i}" for i in range(4)] + [f"neck{i}" for i in range(4)] + \
       [f"fused{i}" for i in range(4)] + ["head_conv1", "depth"]
for (name, variant, tiny, seed, B, H, W, stride) in MODEL_CASES:
    eng = B200Engine.from_hf_model(make_hf_model(variant, seed, tiny), dev, out_dtype=torch.float32)
    x = torch.from_numpy(model_input(seed, B, H, W)).to(dev)
    runs = []
    for r in range(4):
        out = eng(x)
        runs.append({t: eng.tap(t).clone() for t in TAPS})
    torch.cuda.synchronize()
    print("==", name)
    for t in TAPS:
        d = max((runs[r][t] - runs[0][t]).abs().max().item() for r in range(1, 4))
        nan = any(torch.isnan(runs[r][t]).any().item() for r in range(4))
        print(f"   {t:14s} max replay diff {d:.3e}  max {runs[0][t].abs().max().item():.3e} nan={nan}")
    eng.close()
