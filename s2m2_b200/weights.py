"""Parameter tree of the S2M2 stereo model (names + shapes) and synthetic weights.

The names/shapes reproduce the reference's state-dict contract (SURVEY.md §5
"checkpoint / resume": e.g. ``transformer.uformer_list.0.enc_attn0.cross_attn.attn.q.weight (C,C)``,
``refiner.gru.convz1.weight (C,2C,3,1)``, ``upsample_mask_1x.conv_disp.0.weight (1,16,3,3)``) so a
reference checkpoint ``CH{C}NTR{N}.pth`` loads unchanged (reference: s2m2.py:14-78,
model_utils.py:12-43).  ``tests/test_oracle.py`` checks the list against the real reference
module when /root/reference is present.

``random_state_dict`` generates deterministic synthetic weights with PyTorch's default-init
*distribution* (uniform ±1/sqrt(fan_in)); there is no network for real checkpoints.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...]]]

MODEL_TABLE = {  # reference: model_utils.py:12-17
    "S": (128, 1), "M": (192, 2), "L": (256, 3), "XL": (384, 3),
}


def _conv(p, cin, cout, kh, kw=None, bias=True) -> Spec:
    kw = kh if kw is None else kw
    s = [(f"{p}.weight", (cout, cin, kh, kw))]
    if bias:
        s.append((f"{p}.bias", (cout,)))
    return s


def _convT(p, cin, cout, k) -> Spec:
    return [(f"{p}.weight", (cin, cout, k, k)), (f"{p}.bias", (cout,))]


def _lin(p, cin, cout, bias=True) -> Spec:
    s = [(f"{p}.weight", (cout, cin))]
    if bias:
        s.append((f"{p}.bias", (cout,)))
    return s


def _attn(p, dim, heads, pe=False) -> Spec:
    s = _lin(f"{p}.q", dim, dim, False) + _lin(f"{p}.k", dim, dim, False) + _lin(f"{p}.v", dim, dim) \
        + _lin(f"{p}.proj", dim, dim, False)
    if pe:
        s += _lin(f"{p}.pe_proj", 32, dim // heads)
    return s


def _ffn(p, dim) -> Spec:
    return _lin(f"{p}.ffn.0", dim, dim) + _lin(f"{p}.ffn.2", dim, dim)


def _basic_attn(p, dim, heads) -> Spec:       # attentions.py:324-345 registration order
    return _attn(f"{p}.cross_attn.attn", dim, heads) + _attn(f"{p}.self_attn.attn", dim, heads) \
        + _ffn(f"{p}.ffn_c", dim) + _ffn(f"{p}.ffn", dim)


def _global_attn(p, dim, heads, cross, pe) -> Spec:  # attentions.py:284-309
    s = _attn(f"{p}.self_attn.attn", dim, heads, pe)
    if cross:
        s += _attn(f"{p}.cross_attn.attn", dim, heads) + _ffn(f"{p}.ffn_c", dim)
    return s + _ffn(f"{p}.ffn", dim)


def _conv_block(p, dim) -> Spec:
    return _conv(f"{p}.convs.0", dim, dim, 3) + _conv(f"{p}.convs.2", dim, dim, 3) \
        + _conv(f"{p}.convs_1x.0", dim, dim, 1) + _conv(f"{p}.convs_1x.2", dim, dim, 1)


def _fusion(p, dim, k) -> Spec:
    return _conv(f"{p}.feature_gate.0", 2 * dim, dim, k) + _conv(f"{p}.feature_gate.2", dim, dim, 1) \
        + _conv(f"{p}.feature_fusion.0", 2 * dim, 2 * dim, k) + _conv(f"{p}.feature_fusion.2", 2 * dim, dim, 1)


def _updown(p, dims) -> Spec:
    d0, d1, d2 = dims
    return _conv(f"{p}.down_conv0.1", d0, d1, 1) + _conv(f"{p}.down_conv1.1", d1, d2, 1) + _conv(f"{p}.down_conv2.1", d2, d2, 1) \
        + _conv(f"{p}.up_conv0.1", d1, d0, 1) + _conv(f"{p}.up_conv1.1", d2, d1, 1) + _conv(f"{p}.up_conv2.1", d2, d2, 1)


def _unet(p, dims, n_attn, pe) -> Spec:       # unet.py:12-63
    d0, d1, d2 = dims
    s = _updown(p, dims)
    s += _fusion(f"{p}.concat_conv0", d0, 1) + _fusion(f"{p}.concat_conv1", d1, 1) + _fusion(f"{p}.concat_conv2", d2, 1)
    s += _conv_block(f"{p}.enc0", d0) + _conv_block(f"{p}.enc1", d1) + _conv_block(f"{p}.enc2", d2)
    for i in range(n_attn):
        s += _global_attn(f"{p}.enc3s.{i}", d2, 8, False, pe)
    s += _conv_block(f"{p}.dec0", d0) + _conv_block(f"{p}.dec1", d1) + _conv_block(f"{p}.dec2", d2)
    for i in range(n_attn):
        s += _global_attn(f"{p}.dec3s.{i}", d2, 8, False, False)
    return s


def _mrt(p, dims) -> Spec:                    # stacked_MRT.py:11-86
    d0, d1, d2 = dims
    s = _updown(p, dims)
    s += _fusion(f"{p}.down_concat1", d1, 1) + _fusion(f"{p}.down_concat2", d2, 1) + _fusion(f"{p}.down_concat3", d2, 1)
    s += _fusion(f"{p}.up_concat0", d0, 1) + _fusion(f"{p}.up_concat1", d1, 1) + _fusion(f"{p}.up_concat2", d2, 1)
    s += _basic_attn(f"{p}.enc_attn0", d0, 1) + _basic_attn(f"{p}.enc_attn1", d1, 2) + _basic_attn(f"{p}.enc_attn2", d2, 4)
    for i in range(2):
        s += _global_attn(f"{p}.enc_attn3s.{i}", d2, 8, True, False)
    s += _basic_attn(f"{p}.dec_attn0", d0, 1) + _basic_attn(f"{p}.dec_attn1", d1, 2) + _basic_attn(f"{p}.dec_attn2", d2, 4)
    for i in range(2):
        s += _global_attn(f"{p}.dec_attn3s.{i}", d2, 8, True, False)
    return s


def param_spec(C: int, NTR: int) -> Spec:
    """All (name, shape) pairs of ``S2M2(C, 1, NTR).state_dict()`` (s2m2.py:14-67)."""
    s: Spec = []
    b = "cnn_backbone"                          # submodules.py:66-86
    s += _conv(f"{b}.conv0.0", 3, 16, 1) + _conv(f"{b}.conv0.2", 16, 16, 1)
    s += _conv(f"{b}.conv1_down.0", 16, 64, 5) + _conv(f"{b}.conv1_down.2", 64, C, 3)
    s += [(f"{b}.norm1.weight", (C,)), (f"{b}.norm1.bias", (C,))]
    s += _conv(f"{b}.conv2.0", C, C, 3) + _conv(f"{b}.conv2.2", C, C, 3) + _conv(f"{b}.conv2_down.0", C, C, 3)
    s += _unet("feat_pyramid", (C, C, 2 * C), 2 * NTR, True)
    for l in range(NTR):
        s += _mrt(f"transformer.uformer_list.{l}", (C, C, 2 * C))
    s += [("disp_init.layer_norm.weight", (C,)), ("disp_init.layer_norm.bias", (C,))]
    u = "upsample_mask_1x"                      # submodules.py:121-135
    s += _convT(f"{u}.conv_disp.0", 1, 16, 3) + _convT(f"{u}.conv_rgb.0", 3, 16, 3) + _convT(f"{u}.conv_ctx", C, 16, 2)
    s += _conv(f"{u}.conv_concat.0", 48, 48, 3) + _convT(f"{u}.conv_concat.2", 48, 9, 1)
    u = "upsample_mask_4x_refine"               # submodules.py:99-108
    s += _convT(f"{u}.conv_x", C, 64, 2) + _conv(f"{u}.conv_y", C, 64, 3)
    s += _conv(f"{u}.conv_concat.0", 128, 128, 3) + _convT(f"{u}.conv_concat.2", 128, 9, 2)
    g = "global_refiner"                        # refinenet.py:43-57
    s += _conv(f"{g}.init_feat.0", 2 + C, C, 3) + _conv(f"{g}.init_feat.2", C, C, 1)
    s += _unet(f"{g}.refine_unet", (C, C, C), 1, False)
    s += _conv(f"{g}.out_feat.0", C, 1, 3)
    s += _fusion("feat_fusion_layer", C, 3)
    r = "refiner"                               # refinenet.py:80-124
    s += _conv(f"{r}.disp_feat.0", 1, 96, 3) + _conv(f"{r}.disp_feat.2", 96, 96, 3)
    s += _conv(f"{r}.corr_feat1.0", 9, 96, 1) + _conv(f"{r}.corr_feat1.2", 96, 64, 1)
    s += _conv(f"{r}.corr_feat2.0", 9, 96, 1) + _conv(f"{r}.corr_feat2.2", 96, 64, 1)
    s += _conv(f"{r}.conf_occ_feat.0", 2, 64, 3) + _conv(f"{r}.conf_occ_feat.2", 64, 32, 1)
    s += _conv(f"{r}.disp_corr_ctx_cat.0", 256 + C, 2 * C, 1) + _conv(f"{r}.disp_corr_ctx_cat.2", 2 * C, C, 3)
    s += _unet(f"{r}.refine_unet", (C, C, 2 * C), 1, False)
    s += _conv(f"{r}.disp_update.0", C, C, 3) + _conv(f"{r}.disp_update.2", C, 1, 3, bias=False)
    s += _conv(f"{r}.conf_occ_update.0", C, C, 3) + _conv(f"{r}.conf_occ_update.2", C, 2, 3, bias=False)
    for n in ("z1", "r1", "q1"):
        s += _conv(f"{r}.gru.conv{n}", 2 * C, C, 3, 1)
    for n in ("z2", "r2", "q2"):
        s += _conv(f"{r}.gru.conv{n}", 2 * C, C, 1, 3)
    s += _conv("ctx_feat.0", C, C, 1) + _conv("ctx_feat.2", C, C, 1)
    return s


def _fan_in(name: str, shape: Tuple[int, ...]) -> int:
    if len(shape) == 1:
        return 0
    if len(shape) == 2:
        return shape[1]
    # ConvTranspose2d weights are (in, out, kh, kw); torch computes fan_in from dim 1 for them too.
    return shape[1] * shape[2] * shape[3]


def random_state_dict(C: int, NTR: int, seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights, PyTorch default-init distribution (uniform ±1/sqrt(fan_in)).

    Norm weights are 1 and norm biases 0 like torch's defaults.  ``gain`` scales the matrices
    (a gain > 1 yields more decisive matching margins — SURVEY.md §8(d) "Weights").
    """
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    spec = param_spec(C, NTR)
    fan_of = {}
    for name, shape in spec:
        if name.endswith(".weight") and len(shape) > 1:
            fan_of[name[:-7]] = _fan_in(name, shape)
    for name, shape in spec:
        base, leaf = name.rsplit(".", 1)
        if "norm" in base and len(shape) == 1:
            sd[name] = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            continue
        fan = fan_of[base]
        bound = gain / math.sqrt(fan) if leaf == "weight" else 1.0 / math.sqrt(fan)
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def synthetic_pair(B: int, H: int, W: int, seed: int = 1, shift: int = 24, smooth: bool = True):
    """Synthetic stereo pair (SURVEY.md §8(d) "Inputs"): integer-valued fp32 in [0,255], NCHW.

    ``right`` is ``left`` shifted by ``shift`` px (true disparity); ``smooth`` applies a 3x3 box blur
    and re-quantises, which widens the reference's argmax margins.
    """
    g = torch.Generator().manual_seed(seed)
    base = torch.randint(0, 256, (B, 3, H, W + 64), generator=g).float()
    if smooth:
        base = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(base, (1, 1, 1, 1), mode="replicate"), 3, 1)
        base = base.round().clamp(0, 255)
    return base[..., :W].contiguous(), base[..., shift:shift + W].contiguous()
