"""GPU-box diagnostic: StereoPipeline step by step with faulthandler (finds host-side crashes that pytest hides)."""
import faulthandler, os, sys
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from desktop2stereo_b200 import depth
from desktop2stereo_b200.pipeline import StereoPipeline
from oracle.gen_golden import TINY, synth_frame
from oracle.ref_harness import make_hf_model

dev = torch.device("cuda:0")
P = lambda *a: print(*a, flush=True)
depth.init(make_hf_model("Small", 3, TINY), device=dev, depth_resolution=252)
frames = [synth_frame(100 + i, 270, 480, 4) for i in range(12)]
for ema in (True, False):
    P("create", ema)
    pipe = StereoPipeline(depth_slots=4, display_mode="Full-SBS", use_temporal_smooth=ema)
    P("host run")
    got = [r.copy() for r in pipe.run(iter(frames), host=True)]
    P("host ok", got[0].shape, float(got[-1].mean()))
    pipe.reset()
    P("reset ok")
    dv = [torch.from_numpy(f).to(dev) for f in frames]
    out = [r.clone() for r in pipe.run(iter(dv), host=False)]
    torch.cuda.synchronize()
    P("device ok", float(out[-1].float().mean()), np.array_equal(out[0].cpu().numpy(), got[0]))
    pipe.close()
    P("closed")
    del dv, out
    torch.cuda.synchronize()
    P("freed")
P("done")
