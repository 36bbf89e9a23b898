"""Pin oracle/dav2.py (fp32 functional restatement of Depth Anything V2) on HF transformers' own model:
against the committed goldens (tests/golden/model.npz) and against the HF module run live on the same seeded weights."""
import os

import numpy as np
import pytest
import torch

from oracle import dav2
from oracle.gen_golden import MODEL_CASES, model_input
from oracle.ref_harness import make_hf_model


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "model.npz"))


@pytest.mark.parametrize("case", [c for c in MODEL_CASES if c[0] != "tiny_518"], ids=lambda c: c[0])
def test_oracle_matches_hf_and_golden(g, case):
    name, variant, tiny, seed, B, H, W, stride = case
    model = make_hf_model(variant, seed, tiny)
    x = torch.from_numpy(model_input(seed, B, H, W))
    with torch.no_grad():
        ref = model(pixel_values=x).predicted_depth
        out = dav2.forward(model.state_dict(), dav2.cfg_from_hf(model.config), x)
    gold = g[name]
    scale = float(np.abs(gold).max())
    assert scale > 0.5 and float((gold > 0).mean()) > 0.3          # non-degenerate depth
    assert np.abs(ref.numpy()[:, ::stride, ::stride] - gold).max() <= 1e-5 * scale   # goldens reproduce on this box
    assert (out - ref).abs().max().item() <= 2e-5 * scale          # fp32 restatement == HF module


def test_blob_layout_matches_config():
    """pack_state_dict emits exactly the number of floats the engine's reader walks (host-side logic, no GPU)."""
    from desktop2stereo_b200.weights import config_from_hf, pack_state_dict
    from oracle.gen_golden import TINY
    model = make_hf_model("Small", 0, TINY)
    cfg = config_from_hf(model.config)
    blob = pack_state_dict(model.state_dict(), cfg)
    D, L, F, c = cfg.hidden, cfg.layers, cfg.fusion, list(cfg.neck)
    n = D * 588 + D + D + (1 + 37 * 37) * D
    n += L * (2 * D + 3 * D * D + 3 * D + D * D + D + 2 * D + 4 * D * D + 4 * D + 4 * D * D + D) + 2 * D
    n += sum(ci * D + ci for ci in c) + c[0] * c[0] * 16 + c[0] + c[1] * c[1] * 4 + c[1] + c[3] * c[3] * 9 + c[3]
    n += sum(F * ci * 9 for ci in c) + 4 * (F * F + F + 4 * (F * F * 9 + F))
    n += (F // 2) * F * 9 + F // 2 + 32 * (F // 2) * 9 + 32 + 32 + 1
    assert blob.dtype == np.float32 and blob.size == n
    assert [cfg.out_indices[i] for i in range(4)] == [1, 2, 3, 4] and cfg.pos_grid == 37 and cfg.mlp_hidden == 4 * D
