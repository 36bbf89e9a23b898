"""Pins oracle/cpu_pipeline.py (bench.py's `cpu_baseline` / `--impl reference` arm) on the UNMODIFIED reference: the same frame and
seeded weights through /root/reference/depth.py's own process -> predict_depth -> make_sbs on CPU (bf16 autocast, non-CUDA
branches) and through the port must give the same float32 frame ("CPU port unpinned").  Runs where the
reference tree exists (the build container); skipped on the GPU box."""
import os

import numpy as np
import pytest
import torch

from oracle.gen_golden import TINY, synth_frame
from oracle.ref_harness import REFERENCE_ROOT, load_reference, make_hf_model

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("mode", ["Full-SBS", "Half-SBS"])
def test_cpu_port_equals_unmodified_reference(mode):
    from oracle.cpu_pipeline import ReferenceCPUPipeline
    seed, h, w, res = 13, 180, 320, 126
    torch.set_num_threads(1)          # as shipped (depth.py:19); also makes ATen's CPU kernels deterministic
    ref = load_reference("Small", depth_resolution=res, fp16=False, seed=seed, tiny=TINY)   # FP16 off: on CPU the reference cannot mix fp16 tensors with its bf16 autocast
    assert not ref.IS_CUDA
    frames = [synth_frame(seed + i, h, w, 4) for i in range(3)]
    port = ReferenceCPUPipeline(make_hf_model("Small", seed, TINY), res, foreground_scale=ref.FOREGROUND_SCALE, aa_strength=ref.AA_STRENGTH)
    for i, f in enumerate(frames):     # 3 consecutive frames: the DepthStabilizer EMA state is part of the path
        rgb = ref.process(f.copy(), h)
        d = ref.predict_depth(rgb)
        want = ref.make_sbs(rgb, d, ipd_uv=0.064, depth_ratio=2.0, convergence=0.0, display_mode=mode)
        got = port.frame(f, mode, 2.0)
        assert want.dtype == np.float32 and got.shape == want.shape
        assert d.dtype == torch.bfloat16      # CPU depth comes out of bf16 autocast
        err = np.abs(got - want)
        print(mode, "frame", i, "max", err.max(), "mean", err.mean())
        assert np.array_equal(got, want), (mode, i, float(err.max()))
