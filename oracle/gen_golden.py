"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/depth.py) on CPU.

Run in the build container only (the reference tree does not travel to the GPU box):
    python -m oracle.gen_golden [warp] [post] [pre] [model] [e2e] [overlay] [vda]
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


def _cuda_branch_process(depth_mod):
    """The reference defines its CUDA `process` only when IS_CUDA is true at import (depth.py:540-566).
    Compile that exact FunctionDef from the reference source and run it on CPU in the module's namespace."""
    import ast
    src = open(depth_mod.__file__).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.If) and isinstance(node.test, ast.Name) and node.test.id == "IS_CUDA":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "process":
                    ns = dict(depth_mod.__dict__)
                    exec(compile(ast.Module(body=[fn], type_ignores=[]), depth_mod.__file__, "exec"), ns)
                    return ns["process"]
    raise RuntimeError("CUDA-branch process() not found in the reference")


def synth_depth(seed, H, W):
    """Raw-depth-like positive map, regenerable on any box from the seed (tests rebuild it)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    d = 2.0 + np.sin(xx / W * 7.0 + rng.random() * 6) + np.cos(yy / H * 4.0 + rng.random() * 6)
    d += 0.8 * (np.hypot(xx - W * rng.random(), yy - H * rng.random()) < min(H, W) / 4)
    d += 0.05 * rng.standard_normal((H, W)).astype(np.float32)
    return np.maximum(d, 0).astype(np.float32)


def synth_frame(seed, h, w, ch=4):
    """BGRA/BGR u8 frame: smooth gradients + blocks + noise (regenerable from the seed)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([128 + 100 * np.sin(xx / w * (3 + c) + c) * np.cos(yy / h * (2 + c)) for c in range(3)], -1)
    img[h // 4: h // 2, w // 3: 2 * w // 3] += 60
    img += rng.normal(0, 12, img.shape)
    img = np.clip(img, 0, 255).astype(np.uint8)
    if ch == 4:
        img = np.concatenate([img, np.full((h, w, 1), 255, np.uint8)], -1)
    return img


def gen_pre(depth_mod):
    """process() [CUDA branch], _resize_patch_aligned_t [CUDA branch] and the /255, mean/std normalisation."""
    out = {"versions": _versions()}
    proc = _cuda_branch_process(depth_mod)
    depth_mod.IS_CUDA = True
    try:
        cases = [(0, 135, 240, 4, 135), (1, 270, 480, 4, 135), (2, 200, 301, 3, 120)]
        for (seed, h, w, ch, target_h) in cases:
            frame = synth_frame(seed, h, w, ch)
            for dt in (torch.float32, torch.float16):
                if dt == torch.float16 and target_h < h:
                    continue  # ATen's CPU antialias kernel has no Half implementation
                old = depth_mod.DTYPE
                depth_mod.DTYPE = dt
                ns_proc = proc.__globals__
                ns_proc["DTYPE"] = dt
                rgb = proc(frame.copy(), target_h)
                depth_mod.DTYPE = old
                arr = rgb.float().numpy()
                out[f"proc{seed}_{str(dt).split('.')[1]}"] = arr.astype(np.uint8) if target_h >= h else arr
            out[f"proc{seed}_meta"] = np.array([seed, h, w, ch, target_h])
        # resize + normalise: u8 CHW input (fp32 arithmetic) at several aspect ratios, incl. 16:9 -> 294x518-like
        for (seed, h, w, target) in [(10, 135, 240, 70), (11, 240, 135, 70), (12, 224, 224, 98), (13, 270, 480, 126), (14, 98, 98, 98)]:
            frame = synth_frame(seed, h, w, 3)
            t = torch.from_numpy(frame[..., ::-1].copy()).permute(2, 0, 1).unsqueeze(0)  # RGB CHW u8
            r = depth_mod._resize_patch_aligned_t(t, target, 14)
            x = r.to(torch.float32) / 255.0
            x = (x - depth_mod.MEAN) / depth_mod.STD
            out[f"pre{seed}_resized"] = r.float().numpy()[0] if r.dtype != torch.uint8 else r.numpy()[0].astype(np.float32)
            out[f"pre{seed}_input"] = x.numpy()[0]
            out[f"pre{seed}_meta"] = np.array([seed, h, w, target])
        # shapes table (integer maths, bit-exact target)
        shapes = []
        for (h, w) in [(1080, 1920), (2160, 3840), (518, 518), (720, 1280), (1440, 2560), (1200, 1600), (1080, 2560), (333, 777), (14, 14), (100, 37)]:
            for target in (518, 336, 294, 70):
                t = torch.zeros(1, 3, h, w, dtype=torch.uint8)
                r = depth_mod._resize_patch_aligned_t(t, target, 14) if max(h, w) * 3 < 20000 else None
                shapes.append([h, w, target, r.shape[2], r.shape[3]])
        out["shapes"] = np.array(shapes)
    finally:
        depth_mod.IS_CUDA = False
    np.savez_compressed(os.path.join(GOLDEN, "pre.npz"), **out)
    print("pre.npz written")


def gen_post(depth_mod):
    """post_process_depth, DepthStabilizer (3-frame EMA) and the final upsample, in fp32 / bf16 / fp16."""
    out = {"versions": _versions()}
    i = 0
    for (H, W, oh, ow) in [(42, 70, 135, 240), (70, 70, 98, 98), (294, 518, 540, 960)]:
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            big = H > 100
            if big and dt != torch.float32:
                continue
            depth_mod.depth_stabilizer.prev = None
            frames = []
            for f in range(3):
                raw = torch.from_numpy(synth_depth(100 + i * 10 + f, H, W)).to(dt)
                try:
                    pp = depth_mod.post_process_depth(raw.clone())
                except RuntimeError as e:  # an op missing for this dtype on CPU
                    print("skip", dt, e)
                    pp = None
                    break
                st = depth_mod.depth_stabilizer(pp.clone())
                up = torch.nn.functional.interpolate(st[None, None], size=(oh, ow), mode="bilinear", align_corners=False)[0, 0]
                frames.append((pp.float().numpy(), st.float().numpy().copy(), up.float().numpy()))
            if pp is None:
                continue
            sub = 7 if big else 1
            for f, (pp, st, up) in enumerate(frames):
                out[f"post{i}_f{f}_pp"] = pp[::sub, ::sub]
                out[f"post{i}_f{f}_ema"] = st[::sub, ::sub]
                out[f"post{i}_f{f}_up"] = up[::sub * 3, ::sub * 3]
            out[f"post{i}_meta"] = np.array([H, W, oh, ow, {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dt], sub, 100 + i * 10])
            i += 1
    out["n_cases"] = np.array(i)
    out["fg_aa"] = np.array([depth_mod.FOREGROUND_SCALE, depth_mod.AA_STRENGTH], np.float64)
    depth_mod.depth_stabilizer.prev = None
    np.savez_compressed(os.path.join(GOLDEN, "post.npz"), **out)
    print(f"post.npz: {i} cases")


POST_METRIC_CASES = [(0, 42, 70, "float32", 1), (1, 294, 518, "float32", 7), (2, 70, 70, "bfloat16", 1), (3, 3, 3, "float32", 1)]


def synth_metric_depth(seed, H, W):
    """Metric-depth-like map (metres) with invalid (<= 0) holes, regenerable from the seed."""
    rng = np.random.default_rng(seed)
    d = 0.5 + 4.0 * synth_depth(seed, H, W) / 4.0
    d[rng.random((H, W)) < 0.07] = 0.0
    d[H // 5: H // 4, W // 3: W // 2] = -1.0
    return d.astype(np.float32)


def gen_post_metric(depth_mod):
    """post_process_depth with is_metric() true (depth.py:837-841: 1/d on d > 0, percentiles over the valid values only)."""
    out = {"versions": _versions()}
    old = depth_mod._IS_METRIC
    depth_mod._IS_METRIC = True
    try:
        for (seed, H, W, dt, sub) in POST_METRIC_CASES:
            raw = torch.from_numpy(synth_metric_depth(seed, H, W)).to(getattr(torch, dt))
            pp = depth_mod.post_process_depth(raw.clone())
            out[f"m{seed}"] = pp.float().numpy()[::sub, ::sub]
    finally:
        depth_mod._IS_METRIC = old
    np.savez_compressed(os.path.join(GOLDEN, "post_metric.npz"), **out)
    print("post_metric.npz written")


from desktop2stereo_b200.synth import TINY_CFG as TINY  # noqa: E402
MODEL_CASES = [  # (name, variant, tiny cfg, seed, B, H, W, stored stride)
    ("tiny_70x98", "Small", TINY, 3, 2, 70, 98, 1),
    ("tiny_518", "Small", TINY, 4, 1, 518, 518, 7),
    ("small_70x126", "Small", None, 5, 1, 70, 126, 1),
]


def model_input(seed, B, H, W):
    """Normalised-image-like input, regenerable from the seed."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    base = np.stack([np.sin(xx / W * (4 + c) + b) * np.cos(yy / H * (3 + c) - b) for b in range(B) for c in range(3)], 0)
    x = base.reshape(B, 3, H, W) * 1.5 + rng.normal(0, 0.3, (B, 3, H, W))
    return x.astype(np.float32)


def gen_model(depth_mod):
    """predicted_depth of HF's DepthAnythingForDepthEstimation (fp32, no autocast) on seeded weights and inputs."""
    from .ref_harness import make_hf_model
    out = {"versions": _versions()}
    for (name, variant, tiny, seed, B, H, W, stride) in MODEL_CASES:
        model = make_hf_model(variant, seed, tiny)
        x = torch.from_numpy(model_input(seed, B, H, W))
        with torch.no_grad():
            d = model(pixel_values=x).predicted_depth
        out[name] = d.numpy()[:, ::stride, ::stride]
        print(name, tuple(d.shape), float(d.min()), float(d.max()))
    np.savez_compressed(os.path.join(GOLDEN, "model.npz"), **out)


E2E = dict(seed=11, h=90, w=160, depth_resolution=70)


def gen_e2e(_unused):
    """process -> predict_depth -> make_sbs of the reference on one seeded frame, tiny seeded model, fp32 (autocast
    disabled so the golden is a precision reference), CUDA-branch resize semantics, EMA off."""
    import contextlib
    dm = load_reference("Small", depth_resolution=E2E["depth_resolution"], fp16=False, seed=E2E["seed"], tiny=TINY)
    dm.maybe_autocast = lambda *a, **k: contextlib.nullcontext()
    proc = _cuda_branch_process(dm)
    dm.IS_CUDA = True
    try:
        frame = synth_frame(E2E["seed"], E2E["h"], E2E["w"], 4)
        rgb = proc(frame.copy(), E2E["h"])
        depth = dm.predict_depth(rgb, use_temporal_smooth=False)
        out = {"versions": _versions(), "depth": depth.float().numpy()}
        for mode in ("Half-SBS", "Full-SBS"):
            out["sbs_" + mode] = dm.make_sbs(rgb, depth, ipd_uv=0.064, depth_ratio=4.0, convergence=0.0, display_mode=mode).astype(np.float16)
        print("e2e depth", tuple(depth.shape), depth.dtype, float(depth.min()), float(depth.max()), float(depth.mean()))
    finally:
        dm.IS_CUDA = False
    np.savez_compressed(os.path.join(GOLDEN, "e2e.npz"), **out)


OVERLAY_CASES = [  # (seed, H, W, dtype, [fps per call])
    (0, 40, 64, "float32", [59.94]), (1, 130, 200, "float16", [7.0]), (2, 61, 33, "float32", [1234.5]),
    (3, 200, 320, "float32", [float(v) for v in np.linspace(10, 130, 23)]),     # the every-10th-call cache
    (4, 1080, 1920, "float16", [143.7]), (5, 9, 300, "float32", [float("inf")]), (6, 60, 7, "bfloat16", [99.9]),
]


def overlay_rgb(seed, H, W):
    """Low-entropy frame (the fixture compresses well), regenerable from the seed."""
    yy, xx = np.mgrid[0:H, 0:W]
    return np.stack([(yy * (3 + seed) + xx * (5 + c) + 40 * c) % 256 for c in range(3)]).astype(np.float32)


def gen_overlay(depth_mod):
    """overlay_fps with a fresh cache per case; stores the top-left crop that can hold text plus a checksum that the rest
    of the image came back untouched."""
    out = {"versions": _versions()}
    for (seed, H, W, dt, fps_seq) in OVERLAY_CASES:
        rgb = torch.from_numpy(overlay_rgb(seed, H, W)).to(getattr(torch, dt))
        depth_mod._FPS_MASK_CACHE.update(mask=None, frame=0)
        ch, cw = min(H, 64), min(W, 420)
        crops = []
        for fps in fps_seq:
            o = depth_mod.overlay_fps(rgb, fps)
            assert o.dtype == rgb.dtype
            rest_same = torch.equal(o[:, ch:, :], rgb[:, ch:, :]) and torch.equal(o[:, :, cw:], rgb[:, :, cw:])
            assert rest_same
            crops.append(o[:, :ch, :cw].float().numpy().astype(np.uint8))
        out[f"ov{seed}"] = np.stack(crops)
    depth_mod._FPS_MASK_CACHE.update(mask=None, frame=0)
    np.savez_compressed(os.path.join(GOLDEN, "overlay.npz"), **out)
    print("overlay.npz written")


VDA_CASE = dict(encoder="vits", seed=21, H=70, W=98, frames=40, keep=[0, 1, 2, 17, 31, 32, 33, 39])


def vda_frames(seed, T, H, W):
    """T correlated model inputs [T,1,3,H,W] (a drifting pattern + a little per-frame noise), regenerable from the seed."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = []
    for t in range(T):
        base = np.stack([np.sin((xx + 1.5 * t) / W * (4 + c)) * np.cos((yy - 0.7 * t) / H * (3 + c)) for c in range(3)], 0)
        out.append((base * 1.5 + rng.normal(0, 0.1, (3, H, W))).astype(np.float32)[None])
    return np.stack(out)


def gen_vda(_unused):
    """Streaming Video-Depth-Anything: the reference's own module (models/video_depth_anything/vda2_s.py) on CPU in fp32, seeded
    weights (oracle/vda.py make_state_dict), 40 frames so that the 32-frame window wraps.  `fp32=True` disables autocast except for
    one place the reference forces it (dpt_temporal.py:115-118, output_conv2 under autocast => bf16 on CPU); that forced autocast
    is switched off here so that the golden is a precision reference."""
    import contextlib
    import types
    from . import vda
    from .ref_harness import REFERENCE_ROOT
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "easydict" not in sys.modules:      # dpt_temporal.py:19 imports easydict (absent here): attribute-dict stub
        ed = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, **kw):
                super().__init__(**kw)
                self.__dict__ = self
        ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed
    import models.video_depth_anything.dpt_temporal as dt
    dt.maybe_autocast = lambda *a, **k: contextlib.nullcontext()
    from models.video_depth_anything.vda2_s import VideoDepthAnything
    c = VDA_CASE
    enc = vda.ENCODERS[c["encoder"]]
    m = VideoDepthAnything(encoder=c["encoder"], features=enc["features"], out_channels=enc["out_channels"]).eval()
    assert [n for n, _ in vda.param_shapes(c["encoder"])] == list(m.state_dict().keys())
    m.load_state_dict(vda.make_state_dict(c["encoder"], c["seed"]), strict=True)
    frames = vda_frames(c["seed"], c["frames"], c["H"], c["W"])
    out = {"versions": _versions()}
    for t in range(c["frames"]):
        d = m(pixel_values=torch.from_numpy(frames[t]), fp32=True)
        if t in c["keep"]:
            out[f"depth{t}"] = d.numpy()[0, 0]
    print("vda golden", tuple(d.shape), float(d.min()), float(d.max()))
    np.savez_compressed(os.path.join(GOLDEN, "vda.npz"), **out)


def main(argv):
    os.makedirs(GOLDEN, exist_ok=True)
    what = set(argv) or {"warp", "post", "pre", "model", "e2e", "overlay", "vda", "post_metric"}
    depth_mod = load_reference("Small")
    g = globals()
    for name in ["warp", "post", "pre", "model", "e2e", "overlay", "vda", "post_metric"]:
        if name in what and f"gen_{name}" in g:
            g[f"gen_{name}"](depth_mod)


if __name__ == "__main__":
    main(sys.argv[1:])
