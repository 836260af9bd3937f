"""Pre-processing of the reference's wrapper on the device.

``image_pad`` — ``s2m2.core.utils.image_utils.image_pad`` (image_utils.py:27-63), what ``run_stereo_matching`` applies to both
images before the forward (model_utils.py:72-73): centre padding to multiples of 32 with a blurred fill.  Here it is two small
kernels on the GPU (``csrc/match.cu``: pad_pool / pad_fill) that take the uint8 image as it comes from the camera / decoder —
the host never touches a float copy — and ``stereo_match`` chains it with the forward, the fused centre crop
(``S2M2.forward(..., crop_hw=...)``) and the confidence score (``post.confidence_score``), i.e. the whole of
``run_stereo_matching`` (model_utils.py:50-95) without leaving the device.  No CPU fallback.
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


def stereo_match(model, left: torch.Tensor, right: torch.Tensor, margin: int = 100) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``run_stereo_matching`` (model_utils.py:50-95) on the device: pad -> forward -> crop -> interior confidence mean.
    left / right: (B,3,H,W) uint8 / fp32 CUDA images of any size >= 32.  Returns disparity, occlusion, confidence as
    (B,H,W) fp32 maps of the ORIGINAL size and the per-pair confidence score (B,) — everything stays on the GPU."""
    H, W = left.shape[-2:]
    lp, rp = image_pad(left), image_pad(right)
    d, o, c = model(lp, rp, crop_hw=(H, W))
    score = confidence_score(c, None, margin) if min(H, W) > 2 * margin else torch.full((left.shape[0],), float("nan"), device=left.device)
    return d[:, 0], o[:, 0], c[:, 0], score
