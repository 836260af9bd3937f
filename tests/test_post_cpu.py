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


def test_remap_oracle_matches_cv2_and_golden():
    """oracle.post_oracle.remap_u8 (the restatement of OpenCV's fixed-point 8-bit bilinear remap, image_utils.py:108-136) against
    cv2.remap itself where cv2 is importable, and against the committed cv2 outputs (tests/golden/remap_u8.npz) always."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from oracle.post_oracle import remap_u8
    z = np.load(os.path.join(GOLDEN, "remap_u8.npz"))
    for i in range(2):
        assert np.array_equal(remap_u8(z["left"], z[f"c{i}_lx"], z[f"c{i}_ly"]), z[f"c{i}_left"])
        assert np.array_equal(remap_u8(z["right"], z[f"c{i}_rx"], z[f"c{i}_ry"]), z[f"c{i}_right"])
    assert np.array_equal(remap_u8(z["left"], z["syn_mx"], z["syn_my"]), z["syn_out"])
    assert (z["syn_out"] == 0).all(-1).mean() > 0.05           # the synthetic map does leave the image
    try:
        import cv2
    except Exception:
        return
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    mx = rng.uniform(-3, 86, (50, 70)).astype(np.float32)
    my = rng.uniform(-3, 64, (50, 70)).astype(np.float32)
    mx[:5] = np.round(mx[:5] * 64) / 64                         # ties of the 1/32-pixel rounding
    my[:5] = np.round(my[:5])
    assert np.array_equal(remap_u8(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR, cv2.BORDER_CONSTANT))


def test_points_oracle_against_the_reference_get_pointcloud():
    """post_oracle.disparity_to_points against the reference's OWN get_pointcloud (model_utils.py:108-136) run here with open3d
    replaced by a recorder: everything the reference computes itself — the halved intrinsics, depth = baseline*fx/(disp+doffs),
    depth[disp<=0] = 1e9, the float32 cast, depth_scale=1000 / depth_trunc / convert_rgb_to_intensity=False, fx used for both
    axes — is pinned by executing it; the two Open3D calls it then makes (RGBDImage.create_from_color_and_depth ->
    PointCloud.create_from_rgbd_image) are restated from Open3D's documented pinhole back-projection (open3d is not in this
    image: that half stays 'parity unpinned')."""
    import os
    import sys
    import types
    import numpy as np
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present (GPU box)")
    rec = {}
    o3d = types.ModuleType("open3d")

    class Image:
        def __init__(self, a):
            self.a = np.asarray(a)

    class RGBDImage:
        @staticmethod
        def create_from_color_and_depth(color, depth, depth_scale=1000.0, depth_trunc=3.0, convert_rgb_to_intensity=True):
            rec.update(depth=depth.a, depth_scale=depth_scale, depth_trunc=depth_trunc, intensity=convert_rgb_to_intensity)
            return "rgbd"

    class PinholeCameraIntrinsic:
        def __init__(self, w, h, fx, fy, cx, cy):
            rec.update(w=w, h=h, fx=fx, fy=fy, cx=cx, cy=cy)

    class PointCloud:
        @staticmethod
        def create_from_rgbd_image(rgbd, intrinsic):
            return "cloud"
    o3d.geometry = types.SimpleNamespace(Image=Image, RGBDImage=RGBDImage, PointCloud=PointCloud)
    o3d.camera = types.SimpleNamespace(PinholeCameraIntrinsic=PinholeCameraIntrinsic)
    saved = {k: sys.modules.get(k) for k in ("open3d",)}
    sys.modules.setdefault("open3d", o3d)             # only needed for the module's own `import open3d as o3d`
    sys.path.insert(0, "/root/reference/src")
    MU = old_o3d = None
    try:
        from s2m2.core.utils import model_utils as MU
        old_o3d, MU.o3d = MU.o3d, o3d                 # the recorder, whatever an earlier test left in sys.modules
        rng = np.random.default_rng(3)
        H, W = 40, 56
        disp = rng.uniform(-2, 60, (H, W)).astype(np.float32)
        disp[3, 4] = 0.0
        rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        calib = {"cam0": np.array([[1733.74, 0, 792.27], [0, 1733.74, 541.89], [0, 0, 1]]), "baseline": 536.62, "doffs": 0.0}
        assert MU.get_pointcloud(rgb, disp.copy(), calib, depth_trunc=50.0) == "cloud"
    finally:
        if MU is not None and old_o3d is not None:
            MU.o3d = old_o3d
        sys.path.remove("/root/reference/src")
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert rec["depth"].dtype == np.float32 and rec["depth_scale"] == 1000.0 and rec["depth_trunc"] == 50.0 and rec["intensity"] is False
    assert rec["fx"] == rec["fy"] == calib["cam0"][0, 0] / 2.0 and rec["cx"] == calib["cam0"][0, 2] / 2.0 and rec["cy"] == calib["cam0"][1, 2] / 2.0
    # Open3D's part, restated: z = depth / depth_scale, dropped when z <= 0 or beyond depth_trunc; x = (u - cx) z / fx, y = (v - cy) z / fy
    z = rec["depth"].astype(np.float64) / rec["depth_scale"]
    keep = (z > 0) & (z <= rec["depth_trunc"])
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    want = np.stack([(u - rec["cx"]) * z / rec["fx"], (v - rec["cy"]) * z / rec["fy"], z], -1)
    xyz, valid = PO.disparity_to_points(torch.from_numpy(disp)[None], fx=rec["fx"], cx=rec["cx"], cy=rec["cy"], baseline=calib["baseline"],
                                        doffs=calib["doffs"], depth_trunc=50.0)
    assert np.array_equal(valid[0].numpy(), keep)
    assert not keep[3, 4] and not keep[disp <= 0].any()
    got = xyz[0].numpy().astype(np.float64)
    assert np.allclose(got[keep], want[keep], rtol=2e-6, atol=1e-9)
