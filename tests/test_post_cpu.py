"""The post-processing oracle (oracle/post_oracle.py) against the reference's own helpers where they can be imported here."""
import importlib.util
import os

import pytest
import torch

from oracle import post_oracle as PO

REF = "/root/reference/src/s2m2/core/utils/image_utils.py"


def _ref_image_utils():
    if not os.path.exists(REF):
        pytest.skip("reference tree not present")
    src = open(REF).read()
    # the module imports cv2 for file I/O only; image_crop itself is pure torch — load it without cv2
    ns = {}
    import math
    import types
    fake_cv2 = types.ModuleType("cv2")
    import sys
    had = sys.modules.get("cv2")
    sys.modules.setdefault("cv2", fake_cv2)
    try:
        spec = importlib.util.spec_from_file_location("ref_image_utils", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if had is None:
            sys.modules.pop("cv2", None)
    return mod


@pytest.mark.parametrize("hw,crop", [((64, 96), (64, 96)), ((64, 96), (50, 77)), ((96, 128), (95, 100))])
def test_image_crop_matches_reference(hw, crop):
    mod = _ref_image_utils()
    x = torch.randn(2, 1, *hw)
    assert torch.equal(PO.image_crop(x, crop), mod.image_crop(x, crop))


def test_confidence_score_is_the_reference_expression():
    g = torch.Generator().manual_seed(0)
    conf = torch.rand((3, 1, 320, 416), generator=g)
    s = PO.confidence_score(conf, (301, 400), margin=100)
    for b in range(3):
        c = PO.image_crop(conf[b:b + 1], (301, 400)).squeeze().float()       # model_utils.py:88-94 for one pair
        assert abs(float(s[b]) - c[100:-100, 100:-100].mean().item()) < 1e-6


def test_points_oracle_basic_geometry():
    disp = torch.full((1, 4, 6), 8.0)
    disp[0, 0, 0] = 0.0                                  # invalid disparity -> depth 1e9 -> dropped by depth_trunc
    xyz, valid = PO.disparity_to_points(disp, fx=100.0, cx=3.0, cy=2.0, baseline=200.0, doffs=2.0, depth_trunc=10.0)
    z = 200.0 * 100.0 / 10.0 / 1000.0
    assert torch.allclose(xyz[0, 1, 1], torch.tensor([(1 - 3.0) * z / 100.0, (1 - 2.0) * z / 100.0, z]))
    assert not bool(valid[0, 0, 0]) and bool(valid[0, 1, 1]) and float(xyz[0, 0, 0, 2]) == 0.0


@pytest.mark.parametrize("H,W", [(100, 75), (64, 96), (70, 96), (64, 100)])
def test_oracle_image_pad_matches_reference(H, W):
    """post_oracle.image_pad (the checker of the fused GPU pad kernels) against the reference's own image_pad."""
    import os
    import sys
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present (GPU box)")
    sys.path.insert(0, "/root/reference/src")
    from s2m2.core.utils.image_utils import image_pad as ref_pad
    import torch
    from oracle import post_oracle as PO
    g = torch.Generator().manual_seed(H * W)
    img = torch.randint(0, 256, (2, 3, H, W), generator=g, dtype=torch.uint8)
    a = PO.image_pad(img)
    b = ref_pad(img, 32)
    assert a.shape == b.shape
    assert torch.equal(a, b.float())
