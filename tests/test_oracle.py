"""The oracle restatement vs golden vectors produced by the real reference (oracle/make_golden.py)."""
import os
import sys

import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import s2m2_oracle as O
from s2m2_b200.weights import MODEL_TABLE, param_spec, random_state_dict, synthetic_pair


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    cfg, g = load_golden(name)
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"], cfg["gain"])
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    taps = {}
    d, o, c = O.forward(sd, left, right, cfg["NTR"], cfg["pos"], cfg["it"], taps)
    # fp32 restatement of the same ATen ops: differences are summation-order level.
    assert (d - g["disp"]).abs().max() < 2e-4
    assert (o - g["occ"]).abs().max() < 2e-5
    assert (c - g["conf"]).abs().max() < 2e-5
    assert (d - g["disp"]).abs().mean() < 1e-5
    # stage taps
    assert (taps["disp_init.disp"] - g["tap:disp_init#0.0"]).abs().max() < 1e-4
    assert (taps["disp_init.cv"] - g["tap:disp_init#0.3"]).abs().max() < 1e-3
    gr = g["tap:global_refiner#0.0"]          # hook fires before the positivity clamp (s2m2.py:160-161)
    assert (taps["global_refiner"] - (gr.clamp(min=0) if cfg["pos"] else gr)).abs().max() < 1e-4
    for i in range(cfg["it"]):
        ref_occ = g[f"tap:refiner#{i}.occ"]
        # reference hook sees occ before the `occ * occ_mask` glue (s2m2.py:179-180)
        mask = taps[f"refiner.{i}.occ_mask"].float()
        assert (taps[f"refiner.{i}.occ"] - ref_occ * mask).abs().max() < 2e-5


def test_oracle_output_upsample_branch_matches_reference_golden():
    from conftest import GOLDEN_UP2
    cfg, g = load_golden(GOLDEN_UP2)
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"], cfg["gain"])
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    d, o, c = O.forward(sd, left, right, cfg["NTR"], cfg["pos"], cfg["it"], None, True)
    assert d.shape == g["disp"].shape == (cfg["B"], 1, 2 * cfg["H"], 2 * cfg["W"])
    assert (d - g["disp"]).abs().max() < 2e-4 and (o - g["occ"]).abs().max() < 2e-5 and (c - g["conf"]).abs().max() < 2e-5


def test_argmax_is_bit_exact_on_golden():
    """prob_max_ind (submodules.py:226) recovered from the golden disparity: disp = x - corr, integer part."""
    cfg, g = load_golden("tiny_c32_64x96")
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"])
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    taps = {}
    O.forward(sd, left, right, cfg["NTR"], cfg["pos"], cfg["it"], taps)
    ind = taps["disp_init.argmax"]
    assert ind.dtype == torch.int64 and ind.shape == (1, cfg["H"] // 4, cfg["W"] // 4)
    assert int(ind.min()) >= 0 and int(ind.max()) < cfg["W"] // 4
    # positivity: match index never right of the pixel
    xs = torch.arange(cfg["W"] // 4).reshape(1, 1, -1)
    assert bool((ind <= xs).all())


def test_cv_lookup_equals_grid_sample_form():
    torch.manual_seed(0)
    cv = torch.randn(2, 5, 24, 24)
    disp = torch.rand(2, 1, 5, 24) * 30 - 3
    a1, a2 = O.cv_lookup(cv, disp)
    b1, b2 = O.cv_lookup_gridsample(cv, disp)
    assert (a1 - b1).abs().max() < 2e-5 and (a2 - b2).abs().max() < 2e-5


def test_pe_tables_reproduce_get_pe_shape_and_norm():
    pe = O.get_pe(3, 5)
    assert pe.shape == (15, 15, 32)
    assert torch.allclose(pe[..., :16].norm(dim=-1), torch.full((15, 15), 0.5), atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("model_type", ["S", "M", "L", "XL"])
def test_param_spec_matches_reference_state_dict(model_type):
    sys.path.insert(0, "/root/reference/src")
    from s2m2.core.model.s2m2 import S2M2
    C, NTR = MODEL_TABLE[model_type]
    with torch.device("meta"):
        m = S2M2(C, 1, NTR, True, False, 3)
    ref = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert ref == param_spec(C, NTR)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference_other_seed():
    sys.path.insert(0, "/root/reference/src")
    from s2m2.core.model.s2m2 import S2M2
    sd = random_state_dict(32, 1, 11)
    m = S2M2(32, 1, 1, True, False, 2).eval()
    m.load_state_dict(sd, strict=True)
    l, r = synthetic_pair(1, 64, 64, 12)
    with torch.inference_mode():
        ref = m(l, r)
    out = O.forward(sd, l, r, 1, True, 2)
    for a, b in zip(ref, out):
        assert (a - b).abs().max() < 2e-4
