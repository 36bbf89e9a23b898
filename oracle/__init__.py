"""ORACLE — CPU restatements of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import anything from this package.  The product (`desktop2stereo_b200/`) never does.

Pinning status: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so every module here is pinned against outputs of the UNMODIFIED reference run in
the build container (`oracle/gen_golden.py` -> `tests/golden/*.npz`).
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "_build")


def build(force: bool = False) -> str:
    """Compile the C restatements with gcc (strict fp32: -ffp-contract=off). Returns the .so path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    so = os.path.join(BUILD_DIR, "libd2s_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("warp_oracle.c", "dibr_oracle.c", "jpeg_oracle.c")]
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", so, *srcs, "-lm"]
    subprocess.check_call(cmd)
    return so
