"""Pre-processing of the reference's wrapper on the device.

``image_pad`` — ``s2m2.core.utils.image_utils.image_pad`` (image_utils.py:27-63), what ``run_stereo_matching`` applies to both
images before the forward (model_utils.py:72-73): centre padding to multiples of 32 with a blurred fill.  Here it is two small
kernels on the GPU (``csrc/match.cu``: pad_pool / pad_fill) that take the uint8 image as it comes from the camera / decoder —
the host never touches a float copy — and ``stereo_match`` chains it with the forward, the fused centre crop
(``S2M2.forward(..., crop_hw=...)``) and the confidence score (``post.confidence_score``), i.e. the whole of
``run_stereo_matching`` (model_utils.py:50-95) without leaving the device.  ``rectify_images`` is the step in front of it in the
calibration tools (image_utils.py:108-136): cv2.remap of both uint8 images, here one kernel, bit-exact with OpenCV, batched over
candidate rectifications.  No CPU fallback.
"""
import math
from typing import Tuple

import torch

from .post import _ops, confidence_score


def image_pad(img: torch.Tensor, factor: int = 32) -> torch.Tensor:
    """(B,3,H,W) uint8 / fp32 CUDA image -> (B,3,Hn,Wn) fp32, Hn, Wn = the next multiples of ``factor``."""
    if img.dim() != 4 or img.shape[1] != 3 or not img.is_cuda or img.dtype not in (torch.uint8, torch.float32):
        raise ValueError("image_pad expects a (B,3,H,W) uint8 / float32 CUDA tensor")
    B, _, H, W = img.shape
    Hn, Wn = math.ceil(H / factor) * factor, math.ceil(W / factor) * factor
    out = torch.empty((B, 3, Hn, Wn), dtype=torch.float32, device=img.device)
    _ops().image_pad(img.contiguous(), factor, out)
    return out


def rectify_images(left: torch.Tensor, right: torch.Tensor, rectification_data, planar: bool = True):
    """``image_utils.rectify_images`` (image_utils.py:108-136: cv2.remap, INTER_LINEAR, constant border) on the device, bit-exact
    with OpenCV's 8-bit path.  left / right: (H,W,3) uint8 CUDA images as ``read_images`` returns them (image_utils.py:8-24).
    ``rectification_data``: the dict of ``compute_stereo_rectification`` (leftMapX/leftMapY/rightMapX/rightMapY as numpy arrays or
    tensors), or a LIST of such dicts — N candidate rectifications of the same frame, the batch the calibration optimisers
    evaluate one by one (calibration/base.py:15-36, cem.py:64-72).  Returns (N,3,Hd,Wd) uint8 planar tensors — what
    ``S2M2.forward`` / ``stereo_match`` take — or (N,Hd,Wd,3) with ``planar=False``."""
    cands = rectification_data if isinstance(rectification_data, (list, tuple)) else [rectification_data]
    dev = left.device
    if left.dtype != torch.uint8 or right.dtype != torch.uint8 or left.dim() != 3 or left.shape[-1] != 3 or not left.is_cuda:
        raise ValueError("rectify_images expects (H,W,3) uint8 CUDA images")

    def maps(key):
        return torch.stack([torch.as_tensor(c[key], dtype=torch.float32) for c in cands]).to(dev).contiguous()
    out = []
    for img, kx, ky in ((left, "leftMapX", "leftMapY"), (right, "rightMapX", "rightMapY")):
        mx, my = maps(kx), maps(ky)
        N, Hd, Wd = mx.shape
        o = torch.empty((N, 3, Hd, Wd) if planar else (N, Hd, Wd, 3), dtype=torch.uint8, device=dev)
        _ops().remap_u8(img.contiguous(), mx, my, o, planar)
        out.append(o)
    return out[0], out[1]


def stereo_match(model, left: torch.Tensor, right: torch.Tensor, margin: int = 100) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``run_stereo_matching`` (model_utils.py:50-95) on the device: pad -> forward -> crop -> interior confidence mean.
    left / right: (B,3,H,W) uint8 / fp32 CUDA images of any size >= 32.  Returns disparity, occlusion, confidence as
    (B,H,W) fp32 maps of the ORIGINAL size and the per-pair confidence score (B,) — everything stays on the GPU."""
    H, W = left.shape[-2:]
    lp, rp = image_pad(left), image_pad(right)
    d, o, c = model(lp, rp, crop_hw=(H, W))
    score = confidence_score(c, None, margin) if min(H, W) > 2 * margin else torch.full((left.shape[0],), float("nan"), device=left.device)
    return d[:, 0], o[:, 0], c[:, 0], score


def candidate_scores(model, left: torch.Tensor, right: torch.Tensor, rectification_list, margin: int = 100, chunk: int = 8) -> torch.Tensor:
    """The inner loop of the calibration optimisers as one batched device pipeline: for each candidate rectification
    ``evaluate_sample`` (calibration/base.py:15-36) = rectify_images -> compute_confidence_score (model_utils.py:97-107), which
    cem.py:64-72 runs one candidate at a time through host memory.  Here the N candidates are rectified by one kernel, ride the
    batch dimension of the forward ``chunk`` at a time, and only the N scores come back.  left / right: (H,W,3) uint8 CUDA
    images; returns (N,) fp32 scores on the device."""
    l, r = rectify_images(left, right, list(rectification_list))
    out = []
    for i in range(0, l.shape[0], chunk):
        out.append(stereo_match(model, l[i:i + chunk], r[i:i + chunk], margin)[3])
    return torch.cat(out)
