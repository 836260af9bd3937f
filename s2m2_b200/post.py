"""Consumers of the three output maps, fused on the device.

* ``confidence_score``   – the score the online calibration optimises (``calibration/base.py:15-36`` →
  ``model_utils.compute_confidence_score`` → ``run_stereo_matching``, ``model_utils.py:86-95``): centre-crop the confidence map
  back to the un-padded image size and average its interior.  Batched: one value per pair, the maps never leave the GPU.
* ``disparity_to_points`` – ``model_utils.get_pointcloud`` (``model_utils.py:111-136``) without open3d: depth from disparity,
  pin-hole back-projection, trust mask ``conf > .1 and occ > .5`` (``vis_utils.py:62``).

Both enqueue hand-written kernels (``csrc/match.cu``) on the current stream through the C-ABI; there is no CPU fallback.
"""
from typing import Optional, Tuple

import torch

from .ops import CudaOps

_OPS = {}


def _ops() -> CudaOps:
    if "o" not in _OPS:
        _OPS["o"] = CudaOps()
    return _OPS["o"]


def _maps(t: torch.Tensor) -> torch.Tensor:
    """(B,1,H,W) / (B,H,W) / (H,W) fp32 CUDA map -> contiguous (B,H,W)."""
    if t.dim() == 4:
        t = t[:, 0]
    elif t.dim() == 2:
        t = t[None]
    if t.dtype != torch.float32 or not t.is_cuda:
        raise ValueError("expected fp32 CUDA maps")
    return t.contiguous()


def confidence_score(pred_conf: torch.Tensor, img_hw: Optional[Tuple[int, int]] = None, margin: int = 100) -> torch.Tensor:
    """Mean confidence over ``image_crop(pred_conf, img_hw)[margin:-margin, margin:-margin]`` per pair -> (B,) fp32 tensor."""
    c = _maps(pred_conf)
    H, W = c.shape[-2:]
    hw = (H, W) if img_hw is None else (int(img_hw[0]), int(img_hw[1]))
    if margin <= 0:
        raise ValueError("margin must be positive (the reference slices [margin:-margin])")
    out = torch.empty((c.shape[0],), dtype=torch.float32, device=c.device)
    _ops().confidence_score(c, hw, margin, out)
    return out


def disparity_to_points(disp: torch.Tensor, fx: float, cx: float, cy: float, baseline: float, doffs: float = 0.0,
                        conf: Optional[torch.Tensor] = None, occ: Optional[torch.Tensor] = None, depth_scale: float = 1000.0,
                        depth_trunc: float = 1e9):
    """-> (xyz (B,H,W,3) fp32, valid (B,H,W) bool).  ``fx, cx, cy`` are the intrinsics actually used for back-projection
    (``get_pointcloud`` halves the calibration file's values first, model_utils.py:118-120)."""
    d = _maps(disp)
    c = _maps(conf) if conf is not None else None
    o = _maps(occ) if occ is not None else None
    B, H, W = d.shape
    xyz = torch.empty((B, H, W, 3), dtype=torch.float32, device=d.device)
    valid = torch.empty((B, H, W), dtype=torch.uint8, device=d.device)
    _ops().disparity_to_points(d, c, o, fx, cx, cy, baseline, doffs, depth_scale, depth_trunc, xyz, valid)
    return xyz, valid.bool()
