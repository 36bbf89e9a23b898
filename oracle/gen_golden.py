"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/depth.py) on CPU.

Run in the build container only (the reference tree does not travel to the GPU box):
    python -m oracle.gen_golden [warp] [post] [pre] [model] [e2e]
Each fixture stores the seeded inputs, the parameters and the reference's outputs, plus the
torch/transformers versions that produced them.  Test infrastructure only.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from .ref_harness import load_reference

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _versions():
    import transformers
    return np.array([torch.__version__, transformers.__version__])


def _smooth_depth(rng, h, w):
    """Seeded smooth depth in [0,1] with a step edge (SURVEY §8d value distributions)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = 0.5 + 0.25 * np.sin(xx / max(w, 1) * 6.0 + rng.random() * 6) + 0.25 * np.cos(yy / max(h, 1) * 5.0 + rng.random() * 6)
    d[h // 3: 2 * h // 3, w // 4: w // 2] += 0.3
    return np.clip(d, 0, 1).astype(np.float32)


def gen_warp(depth_mod):
    """make_sbs_core on both branches, all display modes, pad on/off, three dtypes; + the SURVEY §8c KAT."""
    rng = np.random.default_rng(1234)
    cases = {}
    i = 0
    for (h, w) in [(4, 16), (11, 24), (36, 64), (30, 41)]:
        big = h > 16
        for mode in ["Full-SBS", "Half-SBS", "Full-TAB", "Half-TAB"]:
            for fill in ([False] if big else [False, True]):
                for dt in ([torch.float32, torch.float16] if big else [torch.float32, torch.float16, torch.bfloat16]):
                    for gather in ([False] if big else [False, True]):
                        conv, ratio = [(0.0, 2.0), (0.5, 4.0)][i % 2]
                        rgb = torch.from_numpy(rng.integers(0, 256, (3, h, w)).astype(np.float32)).to(dt)
                        dep = torch.from_numpy(_smooth_depth(rng, h, w) if i % 3 else rng.random((h, w)).astype(np.float32)).to(dt)
                        depth_mod.IS_DIRECTML = gather
                        out = depth_mod.make_sbs_core(rgb, dep, ipd_uv=0.064, depth_ratio=ratio, display_mode=mode,
                                                      fill_16_9=fill, convergence=conv)
                        key = f"c{i:03d}"
                        cases[key + "_rgb"] = rgb.float().numpy().astype(np.uint8)
                        cases[key + "_depth"] = dep.float().numpy()
                        cases[key + "_out"] = out.float().numpy()
                        cases[key + "_meta"] = np.array([h, w, ["Full-SBS", "Half-SBS", "Full-TAB", "Half-TAB"].index(mode),
                                                         int(fill), {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dt],
                                                         int(gather)], np.int32)
                        cases[key + "_par"] = np.array([0.064, ratio, conv], np.float64)
                        i += 1
    depth_mod.IS_DIRECTML = False
    # KAT (i) of SURVEY §8c: H=4, W=16, depth = linspace(0,1,16) per row
    rgb = torch.arange(16, dtype=torch.float32).view(1, 1, 16).expand(3, 4, 16).contiguous() * 16
    dep = torch.linspace(0, 1, 16).view(1, 16).expand(4, 16).contiguous()
    cases["kat_rgb"] = rgb.numpy()
    cases["kat_depth"] = dep.numpy()
    cases["kat_out_bilinear"] = depth_mod.make_sbs_core(rgb, dep, display_mode="Full-SBS").numpy()
    depth_mod.IS_DIRECTML = True
    cases["kat_out_gather"] = depth_mod.make_sbs_core(rgb, dep, display_mode="Full-SBS").numpy()
    depth_mod.IS_DIRECTML = False
    cases["n_cases"] = np.array(i)
    cases["versions"] = _versions()
    np.savez_compressed(os.path.join(GOLDEN, "warp.npz"), **cases)
    print(f"warp.npz: {i} cases")


def main(argv):
    os.makedirs(GOLDEN, exist_ok=True)
    what = set(argv) or {"warp", "post", "pre", "model", "e2e"}
    depth_mod = load_reference("Small")
    g = globals()
    for name in ["warp", "post", "pre", "model", "e2e"]:
        if name in what and f"gen_{name}" in g:
            g[f"gen_{name}"](depth_mod)


if __name__ == "__main__":
    main(sys.argv[1:])
