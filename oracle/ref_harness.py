"""Import the UNMODIFIED reference `depth.py` on CPU (test infrastructure only).

This file is part of the ORACLE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may use it.  It never ships on the product path.

It follows the recipe recorded in SURVEY.md §8(c):
  1. a scratch CWD holding a `settings.yaml` with the keys `utils.py:819-907` reads,
  2. `sys.path.insert(0, <reference>)`,
  3. `transformers.AutoModelForDepthEstimation.from_pretrained` replaced by a function that
     returns a seeded random-init DepthAnythingForDepthEstimation (no network, no weights on disk),
  4. `import depth`.

`/root/reference` exists only in the build container.  On the GPU box this module is unusable
and `load_reference()` raises; everything that travels uses the committed goldens + `oracle/*.py`.
"""
from __future__ import annotations

import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("D2S_REFERENCE_ROOT", "/root/reference")

from desktop2stereo_b200.synth import DA_V2_VARIANTS, make_hf_model, randomize_constant_params  # noqa: E402,F401  (seeded weights live with the product's bench helpers)


def load_reference(variant: str = "Small", depth_resolution: int = 518, fp16: bool = False,
                   seed: int = 0, display_mode: str = "Half-SBS", tiny: dict | None = None,
                   extra_settings: dict | None = None):
    """Return the reference's `depth` module, imported fresh with a stubbed model loader."""
    import yaml
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree {REFERENCE_ROOT} not present (GPU box?) — use committed goldens")
    settings = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, "settings.yaml")))
    settings.update({
        "Depth Model": f"Depth-Anything-V2-{variant}",
        "Depth Resolution": depth_resolution,
        "FP16": fp16,
        "Display Mode": display_mode,
        "Run Mode": "Legacy Streamer",
        "torch.compile": None, "TensorRT": None, "CoreML": None, "OpenVINO": None, "MIGraphX": None,
        "Language": "EN",
    })
    if extra_settings:
        settings.update(extra_settings)
    work = tempfile.mkdtemp(prefix="d2s_ref_")
    with open(os.path.join(work, "settings.yaml"), "w") as f:
        yaml.safe_dump(settings, f)
    old_cwd = os.getcwd()
    os.chdir(work)
    try:
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        for m in ("depth", "utils"):
            sys.modules.pop(m, None)
        import transformers

        def _stub(*_a, **_k):
            return make_hf_model(variant, seed, tiny)

        transformers.AutoModelForDepthEstimation.from_pretrained = staticmethod(_stub)
        import depth  # noqa: E402  (the reference module)
    finally:
        os.chdir(old_cwd)
    return depth
