"""Build libd2s_b200.so in-tree with nvcc for sm_100a (B200).  No torch types cross this boundary."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libd2s_b200.so")
OBJ_DIR = os.path.join(HERE, "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
# Per-file flags.  The pixel-exact kernels are compiled without FMA contraction: every FMA in them is
# explicit (see warp.cu header).
PER_FILE = {
    "warp.cu": ["-fmad=false"],
    "prepost.cu": ["-fmad=false"],
    "dibr.cu": ["-fmad=false"],
}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "d2s_b200.h"))
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path, *headers]):
            cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(src, []), "-c", path, "-o", obj]
            log = open(obj + ".log", "w")
            procs.append((src, cmd, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log))
    failed = False
    for src, cmd, p, log in procs:
        rc = p.wait()
        log.close()
        text = open(log.name).read()
        if rc != 0:
            failed = True
            sys.stderr.write(f"[d2s build] {src} FAILED\n{' '.join(cmd)}\n{text}\n")
        elif verbose:
            sys.stderr.write(f"[d2s build] {src}\n{text}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
