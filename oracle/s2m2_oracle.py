"""CPU/fp32 oracle for the S2M2 stereo forward path.  TEST INFRASTRUCTURE ONLY.

This file is a *functional* restatement (plain functions over a state-dict, no
nn.Module tree) of the algorithm of junhong-3dv/s2m2's ``S2M2.forward``.  It is
the checker for the CUDA path in ``s2m2_b200``; nothing in the product imports
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.

Pinning: ``tests/golden/*.npz`` were produced by importing the *real* reference
module from /root/reference (script: ``oracle/make_golden.py``) and
``tests/test_oracle.py`` checks this restatement against them, so parity is
pinned to outputs of the reference itself (the reference ships no tests or
golden vectors of its own — SURVEY.md §4).

Every function cites the reference file:line it follows (paths relative to
/root/reference/src/s2m2/core/model/).  The arithmetic itself (conv, linear,
softmax, ...) lives in PyTorch, the reference's un-vendored dependency
(pyproject.toml: unpinned ``torch``; this image: 2.11.0+cu128); we call the
same ATen ops in fp32.

Layout: NCHW like the reference.  ``sd`` is a state-dict with the reference's
key names.  ``taps`` (optional dict) receives named intermediate tensors for
per-stage parity tests.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _conv(sd: SD, name: str, x: Tensor, stride=1, padding=0) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _convT(sd: SD, name: str, x: Tensor, stride=1, padding=0) -> Tensor:
    return F.conv_transpose2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _lin(sd: SD, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(x: Tensor) -> Tensor:
    """LayerNorm over the last dim, eps 1e-5, no affine (attentions.py:117,148,182,213,243)."""
    return F.layer_norm(x, (x.shape[-1],))


def _tap(taps: Optional[dict], key: str, val) -> None:
    if taps is not None:
        taps[key] = val


# --------------------------------------------------------------------------
# attention blocks  (attentions.py)
# --------------------------------------------------------------------------
def _heads(x: Tensor, nh: int) -> Tensor:
    b, n, c = x.shape
    return x.reshape(b, n, nh, c // nh).transpose(1, 2)


def _merge(x: Tensor) -> Tensor:
    b, nh, n, d = x.shape
    return x.transpose(1, 2).reshape(b, n, nh * d)


def self_attn(sd: SD, p: str, x: Tensor, nh: int, pe: Optional[Tensor]) -> Tensor:
    """SelfAttn.forward, attentions.py:33-54.  x: (B', N, C)."""
    q, k, v = (_heads(_lin(sd, f"{p}.{n}", x), nh) for n in ("q", "k", "v"))
    d = q.shape[-1]
    if pe is not None and f"{p}.pe_proj.weight" in sd:
        # explicit softmax + contextual relative PE (attentions.py:42-48)
        score = torch.matmul((d ** -0.5) * q, k.transpose(-1, -2))
        attn = score.softmax(dim=-1)
        out = torch.matmul(attn, v)
        pe_sum = torch.einsum("bnij,ijc->bnic", attn, pe)
        out = out + _lin(sd, f"{p}.pe_proj", pe_sum)
    else:
        out = F.scaled_dot_product_attention(q, k, v)  # attentions.py:50
    return _lin(sd, f"{p}.proj", _merge(out))


def cross_attn(sd: SD, p: str, x: Tensor, y: Tensor, nh: int):
    """CrossAttn.forward (symmetric, shared weights), attentions.py:77-96."""
    qx, kx, vx = (_heads(_lin(sd, f"{p}.{n}", x), nh) for n in ("q", "k", "v"))
    qy, ky, vy = (_heads(_lin(sd, f"{p}.{n}", y), nh) for n in ("q", "k", "v"))
    xo = F.scaled_dot_product_attention(qx, ky, vy)
    yo = F.scaled_dot_product_attention(qy, kx, vx)
    return _lin(sd, f"{p}.proj", _merge(xo)), _lin(sd, f"{p}.proj", _merge(yo))


def ffn(sd: SD, p: str, z: Tensor) -> Tensor:
    """FFN.forward, attentions.py:245-250 (pre-norm, residual on un-normalised input)."""
    h = F.gelu(_lin(sd, f"{p}.ffn.0", _ln(z)))
    return _lin(sd, f"{p}.ffn.2", h) + z


def _cross_block(sd: SD, p: str, z: Tensor, nh: int, seq_shape) -> Tensor:
    """CrossAttnBlock1D/2D.forward, attentions.py:150-161 / 215-226.  z: NHWC (2B,H,W,C)."""
    zn = _ln(z)
    x, y = zn.chunk(2, dim=0)
    shp = x.shape
    x, y = x.reshape(*seq_shape), y.reshape(*seq_shape)
    x, y = cross_attn(sd, f"{p}.attn", x, y, nh)
    return torch.cat([x.reshape(shp), y.reshape(shp)], dim=0) + z


def _self_block(sd: SD, p: str, z: Tensor, nh: int, seq_shape, pe) -> Tensor:
    """SelfAttnBlock1D/2D.forward, attentions.py:119-130 / 185-194."""
    shp = z.shape
    zs = z.reshape(*seq_shape)
    out = self_attn(sd, f"{p}.attn", _ln(zs), nh, pe) + zs
    return out.reshape(shp)


def basic_attn_block(sd: SD, p: str, z: Tensor, nh: int) -> Tensor:
    """BasicAttnBlock.forward (1-D, per scan-line), attentions.py:347-355.  z: NCHW."""
    z = z.permute(0, 2, 3, 1)
    b2, h, w, c = z.shape
    z = _cross_block(sd, f"{p}.cross_attn", z, nh, ((b2 // 2) * h, w, c))
    z = ffn(sd, f"{p}.ffn_c", z)
    z = _self_block(sd, f"{p}.self_attn", z, nh, (b2 * h, w, c), None)
    z = ffn(sd, f"{p}.ffn", z)
    return z.permute(0, 3, 1, 2)


def global_attn_block(sd: SD, p: str, z: Tensor, nh: int, pe: Optional[Tensor]) -> Tensor:
    """GlobalAttnBlock.forward (2-D over all tokens at 1/32), attentions.py:311-321."""
    z = z.permute(0, 2, 3, 1)
    b2, h, w, c = z.shape
    if f"{p}.cross_attn.attn.q.weight" in sd:
        z = _cross_block(sd, f"{p}.cross_attn", z, nh, (b2 // 2, h * w, c))
        z = ffn(sd, f"{p}.ffn_c", z)
    z = _self_block(sd, f"{p}.self_attn", z, nh, (b2, h * w, c), pe)
    z = ffn(sd, f"{p}.ffn", z)
    return z.permute(0, 3, 1, 2).contiguous()


def conv_block(sd: SD, p: str, z: Tensor) -> Tensor:
    """ConvBlock2D.forward, attentions.py:277-281: [3x3 GELU 3x3](z) + [1x1 ReLU 1x1](z)."""
    a = _conv(sd, f"{p}.convs.2", F.gelu(_conv(sd, f"{p}.convs.0", z, padding=1)), padding=1)
    b = _conv(sd, f"{p}.convs_1x.2", F.relu(_conv(sd, f"{p}.convs_1x.0", z)))
    return a + b


def feature_fusion(sd: SD, p: str, z0: Tensor, z1: Tensor) -> Tensor:
    """FeatureFusion.forward, feature_fusion.py:24-33 (kernel size read from the weight)."""
    z = torch.cat([z0, z1], dim=1)
    pad = sd[f"{p}.feature_gate.0.weight"].shape[-1] // 2
    g = _conv(sd, f"{p}.feature_gate.2", F.gelu(_conv(sd, f"{p}.feature_gate.0", z, padding=pad)))
    w = torch.sigmoid(g).clamp(min=0.01, max=0.99)
    f = _conv(sd, f"{p}.feature_fusion.2", F.gelu(_conv(sd, f"{p}.feature_fusion.0", z, padding=pad)))
    return f + w * z0 + (1 - w) * z1


# --------------------------------------------------------------------------
# relative positional encoding (utils.py:5-6, 32-60)
# --------------------------------------------------------------------------
def pe_tables(h: int, w: int, pe_dim: int = 32):
    """The two small tables the (N,N,32) tensor of ``get_pe`` is gathered from.

    Returns (pe_x (2w+1, 16), pe_y (2h+1, 16)); pe[i,j] = .5*cat(pe_x[xi-xj+w-1], pe_y[yi-yj+h-1]).
    Note the reference's sinc uses the literal 3.1415, not math.pi (utils.py:6).
    """
    def table(n: int) -> Tensor:
        L = 2 * n + 1
        sig = 5 / pe_dim
        pos = torch.linspace(-3, 3, L).tanh()
        dim_t = torch.linspace(-1, 1, pe_dim // 2)
        x = (dim_t[None, :] - pos[:, None]) / sig
        s = torch.where(x.abs() < 1e-6, torch.ones_like(x), torch.sin(3.1415 * x) / (3.1415 * x))
        return F.normalize(s, p=2, dim=-1)
    return table(w), table(h)


def get_pe(h: int, w: int, pe_dim: int = 32) -> Tensor:
    """utils.py:32-60 — materialised (hw, hw, 32) tensor (oracle only; the CUDA path never builds it)."""
    px, py = pe_tables(h, w, pe_dim)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    rx = xs[:, None] - xs[None, :] + w - 1
    ry = ys[:, None] - ys[None, :] + h - 1
    return 0.5 * torch.cat([px[rx], py[ry]], dim=2)


# --------------------------------------------------------------------------
# U-Net and MRT  (unet.py, stacked_MRT.py)
# --------------------------------------------------------------------------
def _down(sd: SD, p: str, z: Tensor) -> Tensor:
    return _conv(sd, f"{p}.1", F.avg_pool2d(z, 2))


def _up(sd: SD, p: str, z: Tensor) -> Tensor:
    return _conv(sd, f"{p}.1", F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False))


def unet(sd: SD, p: str, z: Tensor, n_attn: int, use_pe: bool, taps=None):
    """Unet.forward, unet.py:65-112.  Heads at level 3 are always 8 (unet.py:48-63)."""
    pe = None
    if use_pe:
        H, W = z.shape[-2:]
        pe = get_pe(H // 8, W // 8, 32).to(z)
    z0 = conv_block(sd, f"{p}.enc0", z)
    z1 = conv_block(sd, f"{p}.enc1", _down(sd, f"{p}.down_conv0", z0))
    z2 = conv_block(sd, f"{p}.enc2", _down(sd, f"{p}.down_conv1", z1))
    z3 = _down(sd, f"{p}.down_conv2", z2)
    for i in range(n_attn):
        z3 = global_attn_block(sd, f"{p}.enc3s.{i}", z3, 8, pe)
        _tap(taps, f"{p}.enc3s.{i}", z3)
    for i in range(n_attn):
        z3 = global_attn_block(sd, f"{p}.dec3s.{i}", z3, 8, None)  # dec3s ignore pe (unet.py:63,93)
    z2n = conv_block(sd, f"{p}.dec2", feature_fusion(sd, f"{p}.concat_conv2", z2, _up(sd, f"{p}.up_conv2", z3)))
    z1n = conv_block(sd, f"{p}.dec1", feature_fusion(sd, f"{p}.concat_conv1", z1, _up(sd, f"{p}.up_conv1", z2n)))
    z0n = conv_block(sd, f"{p}.dec0", feature_fusion(sd, f"{p}.concat_conv0", z0, _up(sd, f"{p}.up_conv0", z1n)))
    return z0n, z1n, z2n, z3


def mrt(sd: SD, p: str, z0, z1, z2, z3, nh: int = 1):
    """MRT.forward, stacked_MRT.py:89-133."""
    z0 = basic_attn_block(sd, f"{p}.enc_attn0", z0, 1 * nh)
    z1 = feature_fusion(sd, f"{p}.down_concat1", z1, _down(sd, f"{p}.down_conv0", z0))
    z1 = basic_attn_block(sd, f"{p}.enc_attn1", z1, 2 * nh)
    z2 = feature_fusion(sd, f"{p}.down_concat2", z2, _down(sd, f"{p}.down_conv1", z1))
    z2 = basic_attn_block(sd, f"{p}.enc_attn2", z2, 4 * nh)
    z3 = feature_fusion(sd, f"{p}.down_concat3", z3, _down(sd, f"{p}.down_conv2", z2))
    for i in range(2):
        z3 = global_attn_block(sd, f"{p}.enc_attn3s.{i}", z3, 8 * nh, None)
    for i in range(2):
        z3 = global_attn_block(sd, f"{p}.dec_attn3s.{i}", z3, 8 * nh, None)
    z2 = basic_attn_block(sd, f"{p}.dec_attn2", feature_fusion(sd, f"{p}.up_concat2", z2, _up(sd, f"{p}.up_conv2", z3)), 4 * nh)
    z1 = basic_attn_block(sd, f"{p}.dec_attn1", feature_fusion(sd, f"{p}.up_concat1", z1, _up(sd, f"{p}.up_conv1", z2)), 2 * nh)
    z0 = basic_attn_block(sd, f"{p}.dec_attn0", feature_fusion(sd, f"{p}.up_concat0", z0, _up(sd, f"{p}.up_conv0", z1)), 1 * nh)
    return z0, z1, z2, z3


# --------------------------------------------------------------------------
# backbone, matching, refiners, upsampling
# --------------------------------------------------------------------------
def cnn_encoder(sd: SD, p: str, x: Tensor):
    """CNNEncoder.forward, submodules.py:88-93."""
    x = _conv(sd, f"{p}.conv0.2", F.gelu(_conv(sd, f"{p}.conv0.0", x)))
    x = _conv(sd, f"{p}.conv1_down.2", F.gelu(_conv(sd, f"{p}.conv1_down.0", x, stride=2, padding=2)), padding=1)
    x2 = F.group_norm(x, 8, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"])
    x2 = _conv(sd, f"{p}.conv2.2", F.gelu(_conv(sd, f"{p}.conv2.0", x2, padding=1)), padding=1) + x2
    x4 = _conv(sd, f"{p}.conv2_down.0", x2, stride=2, padding=1)
    return x4, x2


def _lse(x: Tensor, dim: int) -> Tensor:
    """logsumexp_stable, submodules.py:147-152 (max-shift, clamp(1e-30) before log)."""
    m = x.max(dim=dim, keepdim=True).values
    y = (x - m).exp().sum(dim=dim, keepdim=True)
    return (m + torch.log(y.clamp(min=1e-30))).squeeze(dim)


def disp_init(sd: SD, p: str, feature: Tensor, use_positivity: bool, ot_iter: int = 3, taps=None):
    """DispInit.forward + _optimal_transport + _sinkhorn, submodules.py:169-243.  feature: (2B,C,h,w)."""
    w = feature.shape[-1]
    f = F.layer_norm(feature.permute(0, 2, 3, 1), (feature.shape[1],),
                     sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"])
    f0, f1 = f.chunk(2, dim=0)
    cv = torch.matmul(f0, f1.transpose(-1, -2))                       # (B,h,w,w)  [i=left x, j=right x]
    mask = torch.triu(torch.ones(w, w, dtype=torch.bool), diagonal=1) if use_positivity \
        else torch.zeros(w, w, dtype=torch.bool)
    mask = mask.to(cv.device)
    a = F.pad(cv.masked_fill(mask, -1e4), (0, 1, 0, 1))               # zero dust-bins (submodules.py:192-193)
    marg = torch.cat([torch.ones(w), torch.tensor([float(w)])]).to(cv.device) / (2 * w)
    log_mu = log_nu = marg.log().reshape(1, 1, w + 1)
    v = log_nu - _lse(a, 2)
    u = log_mu - _lse(a + v.unsqueeze(2), 3)
    for _ in range(ot_iter - 1):
        v = log_nu - _lse(a + u.unsqueeze(3), 2)
        u = log_mu - _lse(a + v.unsqueeze(2), 3)
    out = a + u.unsqueeze(3) + v.unsqueeze(2)
    prob = (out[:, :, :-1, :-1] + math.log(2 * w)).exp().masked_fill(mask, 0)
    ind = prob.argmax(dim=3, keepdim=True)                            # submodules.py:226
    pp = F.pad(prob, (2, 2))
    conf = torch.zeros_like(ind, dtype=prob.dtype)
    corr = torch.zeros_like(conf)
    for k in range(5):                                                # submodules.py:233-236
        wgt = torch.gather(pp, 3, ind + k)
        conf = conf + wgt
        corr = corr + wgt * (ind + k - 2)
    corr = (corr + 1e-4) / (conf + 1e-4)
    xg = torch.arange(w, dtype=cv.dtype, device=cv.device).reshape(1, 1, w)
    disp = (xg - corr.squeeze(3)).unsqueeze(1)
    _tap(taps, "disp_init.argmax", ind.squeeze(3))
    _tap(taps, "disp_init.prob", prob)
    return disp, conf.squeeze(3).unsqueeze(1), prob.sum(dim=3).unsqueeze(1), cv


def cv_lookup(cv: Tensor, disp: Tensor, radius: int = 4):
    """CostVolume.__call__ + bilinear_sampler, submodules.py:7-60, restated as 1-D lerp.

    For pixel (b,h,x): level 1 samples cv[b,h,x,:] at x-disp+dx, level 2 samples the width-pooled
    row at x/2-disp/2+dx, dx=-r..r; zero outside; ``grid_sample`` with the y coordinate on an
    integer row degenerates to linear interpolation along the last axis.
    """
    b, h, w, w2 = cv.shape
    cv2 = F.avg_pool2d(cv.reshape(b * h * w, 1, 1, w2), kernel_size=[1, 2]).reshape(b, h, w, w2 // 2)
    dx = torch.arange(-radius, radius + 1, dtype=cv.dtype, device=cv.device)
    xs = torch.arange(w, dtype=cv.dtype, device=cv.device).reshape(1, 1, w, 1)
    d = disp.reshape(b, h, w, 1)

    def lerp(row: Tensor, pos: Tensor) -> Tensor:
        n = row.shape[-1]
        p0 = torch.floor(pos)
        t = pos - p0
        i0 = p0.long()
        i1 = i0 + 1
        v0 = torch.gather(row, 3, i0.clamp(0, n - 1)) * ((i0 >= 0) & (i0 <= n - 1))
        v1 = torch.gather(row, 3, i1.clamp(0, n - 1)) * ((i1 >= 0) & (i1 <= n - 1))
        return ((1 - t) * v0 + t * v1).permute(0, 3, 1, 2)

    return lerp(cv, xs - d + dx), lerp(cv2, xs / 2 - d / 2 + dx)


def cv_lookup_gridsample(cv: Tensor, disp: Tensor, radius: int = 4):
    """Same as cv_lookup but through F.grid_sample exactly like submodules.py:7-60 (cross-check)."""
    b, h, w, w2 = cv.shape
    dx = torch.linspace(-radius, radius, 2 * radius + 1, dtype=cv.dtype).reshape(1, 1, -1, 1)
    cvf = cv.reshape(b * h * w, 1, 1, w2)
    cv2 = F.avg_pool2d(cvf, kernel_size=[1, 2]).reshape(b * h, 1, w, w2 // 2)
    cv1 = cv.reshape(b * h, 1, w, w2)
    coords = torch.arange(w, dtype=cv.dtype).reshape(1, 1, w, 1).repeat(b, h, 1, 1).reshape(b * h * w, 1, 1, 1)

    def sample(img, x0):
        x0 = x0.reshape(b * h, w, -1, 1)
        y0 = (coords + 0 * dx).reshape(b * h, w, -1, 1)
        Wd, Hd = img.shape[-1], img.shape[-2]
        grid = torch.cat([2 * x0 / (Wd - 1) - 1, 2 * y0 / (Hd - 1) - 1], dim=-1)
        o = F.grid_sample(img, grid, mode="bilinear", align_corners=True)
        return o.reshape(b, h, w, 2 * radius + 1).permute(0, 3, 1, 2)

    dd = disp.reshape(b * h * w, 1, 1, 1)
    return sample(cv1, coords - dd + dx), sample(cv2, coords / 2 - dd / 2 + dx)


def global_refiner(sd: SD, p: str, ctx: Tensor, disp: Tensor, conf: Tensor) -> Tensor:
    """GlobalRefiner.forward, refinenet.py:59-73."""
    mask = 1.0 * (conf > 0.2)
    conf_logit = (mask * conf).logit(eps=1e-1)
    x = torch.cat([disp / 1e2 * mask, conf_logit, ctx], dim=1)
    feat = _conv(sd, f"{p}.init_feat.2", F.gelu(_conv(sd, f"{p}.init_feat.0", x, padding=1)))
    rf = unet(sd, f"{p}.refine_unet", feat, 1, False)[0]
    upd = _conv(sd, f"{p}.out_feat.0", rf, padding=1) * 1e2
    return mask * disp + (1 - mask) * upd


def conv_gru(sd: SD, p: str, h: Tensor, x: Tensor) -> Tensor:
    """ConvGRU.forward, refinenet.py:22-36 (vertical 3x1 pass then horizontal 1x3 pass)."""
    for sfx, pad in (("1", (1, 0)), ("2", (0, 1))):
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(_conv(sd, f"{p}.convz{sfx}", hx, padding=pad))
        r = torch.sigmoid(_conv(sd, f"{p}.convr{sfx}", hx, padding=pad))
        q = torch.tanh(_conv(sd, f"{p}.convq{sfx}", torch.cat([r * h, x], dim=1), padding=pad))
        h = (1 - z) * h + z * q
    return h


def local_refiner(sd: SD, p: str, hidden, ctx, disp, conf, occ, cv, taps=None):
    """LocalRefiner.forward, refinenet.py:126-154."""
    conf_logit = conf.logit(eps=1e-2)
    occ_logit = occ.logit(eps=1e-2)
    corr1, corr2 = cv_lookup(cv, disp, 4)
    cf1 = _conv(sd, f"{p}.corr_feat1.2", F.gelu(_conv(sd, f"{p}.corr_feat1.0", corr1 / 16)))
    cf2 = _conv(sd, f"{p}.corr_feat2.2", F.gelu(_conv(sd, f"{p}.corr_feat2.0", corr2 / 16)))
    df = _conv(sd, f"{p}.disp_feat.2", F.gelu(_conv(sd, f"{p}.disp_feat.0", disp / 1e2, padding=1)), padding=1)
    cof = _conv(sd, f"{p}.conf_occ_feat.2",
                F.gelu(_conv(sd, f"{p}.conf_occ_feat.0", torch.cat([conf_logit, occ_logit], dim=1), padding=1)))
    x = torch.cat([df, cf1, cf2, ctx, cof], dim=1)
    x = _conv(sd, f"{p}.disp_corr_ctx_cat.2", F.gelu(_conv(sd, f"{p}.disp_corr_ctx_cat.0", x)), padding=1)
    rf = unet(sd, f"{p}.refine_unet", x, 1, False)[0]
    hid = conv_gru(sd, f"{p}.gru", hidden, rf)
    dd = _conv(sd, f"{p}.disp_update.2", F.gelu(_conv(sd, f"{p}.disp_update.0", hid, padding=1)), padding=1)
    co = _conv(sd, f"{p}.conf_occ_update.2", F.gelu(_conv(sd, f"{p}.conf_occ_update.0", hid, padding=1)), padding=1)
    cu, ou = co.chunk(2, dim=1)
    return hid, disp + dd, torch.sigmoid(cu + conf_logit), torch.sigmoid(ou + occ_logit)


def unfold3x3(x: Tensor) -> Tensor:
    """custom_unfold, utils.py:9-20: replicate-pad 3x3 neighbourhood, row-major (dy,dx) order."""
    B, C, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
    return torch.cat([xp[:, :, i:i + H, j:j + W] for i in range(3) for j in range(3)], dim=1)


def upsample4x(x: Tensor, logits: Tensor) -> Tensor:
    """S2M2.upsample4x, s2m2.py:101-114 (RAFT convex upsampling)."""
    b, c, h, w = x.shape
    xu = F.interpolate(unfold3x3(x), (h * 4, w * 4), mode="nearest")
    return (xu * logits.softmax(dim=1)).sum(1, keepdim=True)


def filter1x(x: Tensor, logits: Tensor, output_upsample: bool = False) -> Tensor:
    """S2M2.upsample1x, s2m2.py:116-133.  output_upsample=True: neighbourhood nearest x2, logits bilinear x2 (s2m2.py:123-127)."""
    xu = unfold3x3(x)
    if output_upsample:
        xu = F.interpolate(xu, scale_factor=2, mode="nearest")
        logits = F.interpolate(logits, scale_factor=2, mode="bilinear", align_corners=False)
    return (xu * logits.softmax(dim=1)).sum(1, keepdim=True)


def upsample_mask_4x(sd: SD, p: str, hidden: Tensor, feat2x: Tensor) -> Tensor:
    """UpsampleMask4x.forward, submodules.py:110-115."""
    fx = _convT(sd, f"{p}.conv_x", hidden, stride=2)
    fy = _conv(sd, f"{p}.conv_y", feat2x, padding=1)
    t = F.relu(_conv(sd, f"{p}.conv_concat.0", torch.cat([fx, fy], dim=1), padding=1))
    return _convT(sd, f"{p}.conv_concat.2", t, stride=2)


def upsample_mask_1x(sd: SD, p: str, disp: Tensor, rgb: Tensor, ctx: Tensor) -> Tensor:
    """UpsampleMask1x.forward, submodules.py:137-145."""
    fd = F.relu(_convT(sd, f"{p}.conv_disp.0", disp, padding=1))
    fr = F.relu(_convT(sd, f"{p}.conv_rgb.0", rgb, padding=1))
    fc = _convT(sd, f"{p}.conv_ctx", ctx, stride=2)
    t = F.relu(_conv(sd, f"{p}.conv_concat.0", torch.cat([fd, fr, fc], dim=1), padding=1))
    return _convT(sd, f"{p}.conv_concat.2", t)


# --------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------
@torch.no_grad()
def forward(sd: SD, img0: Tensor, img1: Tensor, num_transformer: int, use_positivity: bool = True,
            refine_iter: int = 3, taps: Optional[dict] = None, output_upsample: bool = False):
    """S2M2.forward, s2m2.py:136-197.  img0/img1: (B,3,H,W) fp32 in [0,255], H,W multiples of 32.

    Returns (disparity, occlusion, confidence), each (B,1,H,W) fp32.
    """
    i0 = (img0 / 255.0 - 0.5) * 2
    i1 = (img1 / 255.0 - 0.5) * 2
    f4, f2 = cnn_encoder(sd, "cnn_backbone", torch.cat([i0, i1], dim=0))
    _tap(taps, "cnn_backbone.x4", f4); _tap(taps, "cnn_backbone.x2", f2)
    f2_0 = f2.chunk(2, dim=0)[0]
    py = unet(sd, "feat_pyramid", f4, 2 * num_transformer, True, taps)
    for i, t in enumerate(py):
        _tap(taps, f"feat_pyramid.{i}", t)
    z = py
    for l in range(num_transformer):
        z = mrt(sd, f"transformer.uformer_list.{l}", *z)
    ftr = z[0].contiguous()
    _tap(taps, "transformer", ftr)
    disp, conf, occ, cv = disp_init(sd, "disp_init", ftr, use_positivity, 3, taps)
    _tap(taps, "disp_init.disp", disp); _tap(taps, "disp_init.conf", conf)
    _tap(taps, "disp_init.occ", occ); _tap(taps, "disp_init.cv", cv)
    ftr0 = ftr.chunk(2, dim=0)[0]
    fpy0 = py[0].chunk(2, dim=0)[0]
    disp = global_refiner(sd, "global_refiner", ftr0.contiguous(), disp, conf)
    if use_positivity:
        disp = disp.clamp(min=0)
    _tap(taps, "global_refiner", disp)
    fus = feature_fusion(sd, "feat_fusion_layer", ftr0, fpy0)
    ctx0 = _conv(sd, "ctx_feat.2", F.gelu(_conv(sd, "ctx_feat.0", fus)))
    hidden = torch.tanh(ctx0)
    _tap(taps, "ctx0", ctx0)
    w = fus.shape[-1]
    xs = torch.arange(w, dtype=torch.float32, device=fus.device).reshape(1, 1, 1, -1)
    for it in range(refine_iter):
        hidden, disp, conf, occ = local_refiner(sd, "refiner", hidden, ctx0, disp, conf, occ, cv, taps)
        if use_positivity:
            disp = disp.clamp(min=0)
        occ_mask = torch.ge(xs - disp, 0)
        occ = occ * occ_mask
        _tap(taps, f"refiner.{it}.disp", disp); _tap(taps, f"refiner.{it}.conf", conf)
        _tap(taps, f"refiner.{it}.occ", occ); _tap(taps, f"refiner.{it}.occ_mask", occ_mask)
        _tap(taps, f"refiner.{it}.hidden", hidden)
    m4 = upsample_mask_4x(sd, "upsample_mask_4x_refine", hidden, f2_0)
    disp_up, occ_up, conf_up = upsample4x(disp * 4, m4), upsample4x(occ, m4), upsample4x(conf, m4)
    _tap(taps, "up4x.disp", disp_up)
    m1 = upsample_mask_1x(sd, "upsample_mask_1x", disp_up, i0, f2_0)
    d = filter1x(disp_up, m1, output_upsample)
    if output_upsample:
        d = 2 * d                                                     # s2m2.py:194-195
    return d, filter1x(occ_up, m1, output_upsample), filter1x(conf_up, m1, output_upsample)
