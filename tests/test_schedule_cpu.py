"""The product's kernel schedule (s2m2_b200/model.py) run with the plain-torch op restatement on CPU.

Validates weight packing / fusion / buffer aliasing / op order of the host side against the oracle and
the reference-generated golden vectors.  The CUDA kernels themselves are checked in the -m gpu tests.
"""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import s2m2_oracle as O
from s2m2_b200.model import S2M2
from s2m2_b200.weights import random_state_dict, synthetic_pair
from torch_ops import TorchOps


def _build(cfg, precision="fp32"):
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"], cfg["gain"])
    m = S2M2(cfg["C"], 1, cfg["NTR"], use_positivity=cfg["pos"], refine_iter=cfg["it"], precision=precision,
             output_upsample=cfg.get("up2", False)).eval()
    m.my_load_state_dict(sd)
    m.ops = TorchOps()
    return m, sd


@pytest.mark.parametrize("name", GOLDEN_CASES[:3])
def test_schedule_matches_golden(name):
    cfg, g = load_golden(name)
    m, sd = _build(cfg)
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    m.taps = {}
    d, o, c = m(left, right)
    assert d.shape == (cfg["B"], 1, cfg["H"], cfg["W"]) and d.dtype == torch.float32
    assert (m.taps["disp_init.cv"] - g["tap:disp_init#0.3"]).abs().max() < 2e-3
    if not cfg["pos"]:
        # un-masked matching with random-init weights has exact near-ties: an arg-max flip is legitimate iff the
        # reference's own top-1/top-2 margin there is at rounding level (see tests/test_parity_gpu.py)
        ot = {}
        O.forward(sd, left, right, cfg["NTR"], cfg["pos"], cfg["it"], ot)
        top2 = ot["disp_init.prob"].topk(2, dim=3).values
        margin = (top2[..., 0] - top2[..., 1]) / top2[..., 0].clamp(min=1e-30)
        flips = m.taps["disp_init.argmax"] != ot["disp_init.argmax"]
        assert flips.float().mean() < 0.02 and ((not flips.any()) or float(margin[flips].max()) < 1e-4)
        assert (d - g["disp"]).abs().median() < 1e-4
        return
    assert (m.taps["disp_init.disp"] - g["tap:disp_init#0.0"]).abs().max() < 1e-3
    assert (d - g["disp"]).abs().max() < 2e-3
    assert (d - g["disp"]).abs().mean() < 5e-5
    assert (o - g["occ"]).abs().max() < 1e-4
    assert (c - g["conf"]).abs().max() < 1e-4


def test_schedule_output_upsample_branch_matches_golden():
    """output_upsample=True (s2m2.py:123-127,194-195): (B,1,2H,2W) outputs, disparity doubled."""
    from conftest import GOLDEN_UP2
    cfg, g = load_golden(GOLDEN_UP2)
    m, sd = _build(cfg)
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    d, o, c = m(left, right)
    assert d.shape == (cfg["B"], 1, 2 * cfg["H"], 2 * cfg["W"])
    assert (d - g["disp"]).abs().max() < 5e-3 and (d - g["disp"]).abs().mean() < 2e-4
    assert (o - g["occ"]).abs().max() < 1e-4 and (c - g["conf"]).abs().max() < 1e-4


def test_schedule_stage_taps_match_oracle():
    cfg, g = load_golden("tiny_c32_64x96")
    m, sd = _build(cfg)
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    m.taps = {}
    m(left, right)
    ot = {}
    O.forward(sd, left, right, cfg["NTR"], cfg["pos"], cfg["it"], ot)
    for k in ("cnn_backbone.x4", "cnn_backbone.x2", "feat_pyramid.0", "feat_pyramid.3", "transformer", "ctx0",
              "global_refiner", "refiner.0.disp", "refiner.2.occ", "up4x.disp"):
        err = (m.taps[k] - ot[k]).abs().max().item()
        assert err < 5e-4 * max(1.0, ot[k].abs().max().item()), (k, err)
    assert torch.equal(m.taps["disp_init.argmax"], ot["disp_init.argmax"])


def test_state_dict_surface_matches_reference_contract():
    m = S2M2(32, 1, 1, use_positivity=True, refine_iter=1, precision="fp32")
    sd = random_state_dict(32, 1, 2)
    bad = dict(sd)
    bad["ctx_feat.0.weight"] = torch.zeros(3, 3)          # wrong shape: skipped with a message (s2m2.py:69-78)
    m.my_load_state_dict(bad)
    out = m.state_dict()
    assert list(out.keys()) == list(sd.keys())
    assert torch.equal(out["ctx_feat.2.weight"], sd["ctx_feat.2.weight"])
    assert float(out["ctx_feat.0.weight"].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        m.ops = TorchOps()
        m(torch.zeros(1, 3, 60, 64), torch.zeros(1, 3, 60, 64))


def test_schedule_fp16_packing_matches_golden():
    """precision="fp16" packs differently (zero-padded channels, space-to-depth stem + repacked 5x5 stride-2 conv, fused heads):
    the same schedule on CPU with fp16 storage must stay within fp16 rounding of the fp32 golden."""
    cfg, g = load_golden(GOLDEN_CASES[0])
    m, _ = _build(cfg, precision="fp16")
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    d, o, c = m(left, right)
    assert d.dtype == torch.float32 and torch.isfinite(d).all()
    assert (d - g["disp"]).abs().mean() < 2e-3, float((d - g["disp"]).abs().mean())
    assert (o - g["occ"]).abs().mean() < 1e-3 and (c - g["conf"]).abs().mean() < 1e-3


def test_space_to_depth_repack_of_conv1_down_is_exact():
    """model._pack_d0: Conv2d(16, 64, 5, stride 2, pad 2) == 3x3 stride-1 pad-1 conv over the 2x2 space-to-depth input with the
    repacked weights (taps that fall outside the 5x5 window are zero)."""
    import torch.nn.functional as F
    m = S2M2(32, 1, 1, use_positivity=True, refine_iter=1, precision="fp16")
    m.my_load_state_dict(random_state_dict(32, 1, 3))
    m._ensure_packed(torch.device("cpu"))
    d0 = m._packed["bb"]["d0"]
    assert (d0.kh, d0.kw, d0.stride, d0.cin, d0.cout) == (3, 3, 1, 64, 64)
    w5 = m._w("cnn_backbone.conv1_down.0.weight").half().float()          # the packed weights are the fp16-rounded originals
    b5 = m._w("cnn_backbone.conv1_down.0.bias").float()
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 16, 12, 20), generator=g)
    ref = F.conv2d(x, w5, b5, stride=2, padding=2)
    n, c, H, W = x.shape
    xs = x.permute(0, 2, 3, 1).reshape(n, H // 2, 2, W // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, H // 2, W // 2, 4 * c)
    w3 = d0.w.float().reshape(64, 3, 3, 64).permute(0, 3, 1, 2)
    got = F.conv2d(xs.permute(0, 3, 1, 2), w3, d0.b, stride=1, padding=1)
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
