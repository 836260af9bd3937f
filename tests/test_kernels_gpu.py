"""Each CUDA kernel (through the C-ABI) vs the plain-torch restatement of the same op (tests/torch_ops.py).

Integer results (arg-max index) must be identical; floating-point results within the stated tolerance.
"""
import math

import pytest
import torch

from s2m2_b200.model import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ConvW
from torch_ops import TorchOps

pytestmark = pytest.mark.gpu
DEV = "cuda"
# the torch restatement is the fp32 checker: no TF32 in its convolutions / matmuls
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    from s2m2_b200.ops import CudaOps
    return CudaOps(), TorchOps()


def _rand(shape, dtype, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cpu") * scale).to(DEV).to(dtype)


def _tol(dtype, k=1):
    # fp32: summation-order differences over K terms; fp16: output rounding (2^-11 rel) dominates
    return (2e-5 * max(1.0, math.sqrt(k) / 8), 1e-5) if dtype == torch.float32 else (6e-3, 6e-3)


def _close(a, b, rtol, atol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    assert torch.isfinite(a).all(), what
    assert bool((err <= bound).all()), f"{what}: max err {err.max().item():.3e} (ref max {b.abs().max().item():.3e})"


CONV_CASES = [
    # N, H, W, Cin, Cout, kh, kw, stride, act, res, x2split, shuffle, ld_extra
    (2, 9, 13, 3, 16, 1, 1, 1, ACT_GELU, False, 0, False, 0),        # stem 1x1, scalar path
    (2, 12, 20, 16, 64, 5, 5, 2, ACT_GELU, False, 0, False, 0),      # 5x5 stride 2
    (1, 10, 19, 64, 32, 3, 3, 1, ACT_NONE, True, 0, False, 8),       # 3x3 + residual, strided views
    (2, 8, 12, 32, 32, 3, 3, 2, ACT_NONE, False, 0, False, 0),       # 3x3 stride 2
    (2, 6, 10, 64, 96, 1, 1, 1, ACT_GELU, False, 32, False, 0),      # two-source 1x1 (cat)
    (1, 7, 9, 64, 48, 3, 3, 1, ACT_SIGMOID, False, 32, False, 4),    # two-source 3x3 (feat_fusion_layer)
    (1, 8, 11, 64, 32, 3, 1, 1, ACT_TANH, False, 0, False, 0),       # GRU 3x1
    (1, 8, 11, 64, 32, 1, 3, 1, ACT_SIGMOID, False, 0, False, 0),    # GRU 1x3
    (1, 16, 24, 32, 1, 3, 3, 1, ACT_NONE, False, 0, False, 0),       # head -> 1 channel
    (1, 16, 24, 32, 2, 3, 3, 1, ACT_NONE, False, 0, False, 0),       # head -> 2 channels
    (1, 16, 24, 9, 96, 1, 1, 1, ACT_GELU, False, 0, False, 0),       # corr encoder 9 -> 96
    (1, 16, 24, 1, 96, 3, 3, 1, ACT_GELU, False, 0, False, 3),       # disp encoder from a strided scalar map
    (2, 5, 7, 32, 4 * 16, 1, 1, 1, ACT_NONE, False, 0, True, 0),     # convT 2x2 s2 (Cr=16)
    (1, 6, 5, 128, 4 * 9, 1, 1, 1, ACT_NONE, False, 0, True, 0),     # convT 2x2 s2 (Cr=9, scalar scatter)
    (1, 4, 300, 256, 256, 1, 1, 1, ACT_RELU, True, 0, False, 0),     # wide linear + residual
    (1, 1, 50, 40, 24, 1, 1, 1, ACT_NONE, False, 0, False, 0),       # Cout not multiple of tile
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d(ops, case, dtype):
    cu, th = ops
    N, H, W, Cin, Cout, kh, kw, stride, act, use_res, split, shuffle, ldx = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    ph, pw = kh // 2, kw // 2
    K = kh * kw * Cin
    w = _rand((Cout, K), dtype, g, 1.0 / math.sqrt(K))
    b = _rand((Cout,), torch.float32, g, 0.5)
    cw = ConvW(w, b, kh, kw, stride, ph, pw, Cin, Cout, shuffle, "t")
    C1 = Cin - split if split else Cin
    xbuf = _rand((N, H, W, C1 + ldx), dtype, g)
    x = xbuf[..., :C1] if ldx else xbuf
    x2 = _rand((N, H, W, split), dtype, g) if split else None
    Ho, Wo = (H + 2 * ph - kh) // stride + 1, (W + 2 * pw - kw) // stride + 1
    oshape = (N, 2 * Ho, 2 * Wo, Cout // 4) if shuffle else (N, Ho, Wo, Cout)
    res = _rand(oshape, dtype, g) if use_res else None
    out_a = torch.zeros(oshape[:3] + (oshape[3] + ldx,), dtype=dtype, device=DEV)[..., :oshape[3]] if ldx % 4 == 0 and ldx else \
        torch.zeros(oshape, dtype=dtype, device=DEV)
    out_b = torch.zeros(oshape, dtype=dtype, device=DEV)
    cu.conv2d(x, cw, act, out_a, res, x2)
    th.conv2d(x, cw, act, out_b, res, x2)
    torch.cuda.synchronize()
    rt, at = _tol(dtype, K)
    _close(out_a, out_b, rt, at * 4, f"conv {case}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_conv2d_mixed_dtypes_and_inplace_residual(ops, dtype):
    """fp32 scalar-map input with activation-dtype weights, fp32 output; and out aliasing res (z += f(z))."""
    cu, th = ops
    g = torch.Generator().manual_seed(5)
    x = _rand((1, 12, 18, 4), torch.float32, g)[..., 1:3]
    cw = ConvW(_rand((64, 18), dtype, g, 0.3), _rand((64,), torch.float32, g), 3, 3, 1, 1, 1, 2, 64, False, "t")
    a = torch.zeros((1, 12, 18, 64), dtype=torch.float32, device=DEV)
    b = torch.zeros_like(a)
    cu.conv2d(x, cw, ACT_GELU, a)
    th.conv2d(x, cw, ACT_GELU, b)
    _close(a, b, 1e-4, 1e-4, "mixed")
    z = _rand((1, 6, 40, 64), dtype, g)
    cw2 = ConvW(_rand((64, 64), dtype, g, 0.1), None, 1, 1, 1, 0, 0, 64, 64, False, "t")
    za, zb = z.clone(), z.clone()
    cu.conv2d(z, cw2, ACT_NONE, za, res=za)
    th.conv2d(z, cw2, ACT_NONE, zb, res=zb)
    rt, at = _tol(dtype, 64)
    _close(za, zb, rt, at * 4, "inplace")


def test_row_corr(ops):
    cu, th = ops
    g = torch.Generator().manual_seed(1)
    f = _rand((4, 5, 37, 96), torch.float32, g)
    a = torch.zeros((2, 5, 37, 37), device=DEV)
    b = torch.zeros_like(a)
    cu.row_corr(f, a)
    th.row_corr(f, b)
    _close(a, b, 2e-5, 2e-4, "row_corr")


ATTN_CASES = [
    # S, N, heads, d, cross, pe(h,w)
    (6, 40, 1, 32, False, None),
    (8, 76, 2, 64, True, None),
    (4, 33, 4, 16, True, None),
    (2, 6 * 7, 8, 8, False, (6, 7)),
    (2, 5 * 9, 8, 32, False, (5, 9)),
    (2, 15 * 20, 8, 32, False, (15, 20)),       # S@640x480 pyramid PE block
    (2, 32 * 38, 8, 64, False, (32, 38)),       # L@1216x1024 pyramid PE block
    (2, 20 * 76, 4, 96, False, (20, 76)),       # XL-style wide token grid (76 columns -> 128-key tiles)
    (2, 100, 8, 4, False, None),
    (3, 70, 1, 256, False, None),
    (2, 50, 2, 96, True, None),
    (16, 304, 1, 256, True, None),      # L level-0 scan-line cross attention (tcgen05 whole-row kernel)
    (6, 304, 1, 256, False, None),
    (8, 152, 2, 128, True, None),       # L level 1
    (8, 76, 4, 128, False, None),       # L level 2
    (8, 160, 1, 128, True, None),       # S level 0
    (4, 80, 2, 64, False, None),        # S level 1
    (2, 384, 1, 64, False, None),       # longest row the whole-row kernel takes
    (2, 100, 1, 384, True, None),       # XL level 0 head dim (value channels split over two CTAs)
    (4, 32 * 38, 8, 64, True, None),    # L@1216x1024 2-D global cross attention at 1/32 (mma.sync flash kernel, 19 key tiles)
    (2, 1000, 2, 64, False, None),      # ragged query and key tiles
    (2, 385, 1, 64, False, None),       # just above the whole-row kernel's limit
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("case", ATTN_CASES)
def test_attention(ops, case, dtype):
    cu, th = ops
    S, N, heads, d, cross, pe_hw = case
    C = heads * d
    g = torch.Generator().manual_seed(7)
    qkv = _rand((S, 1, N, 3 * C), dtype, g)
    oa = torch.zeros((S, 1, N, C), dtype=dtype, device=DEV)
    ob = torch.zeros_like(oa)
    pe = pa = pb = None
    if pe_hw:
        pe_c, pe_t = cu.pe_tables(pe_hw[0], pe_hw[1], torch.device(DEV)), th.pe_tables(pe_hw[0], pe_hw[1], torch.device(DEV))
        _close(pe_c[0], pe_t[0], 1e-5, 1e-6, "pe_x table")
        _close(pe_c[1], pe_t[1], 1e-5, 1e-6, "pe_y table")
        pa = torch.zeros((S, N, heads, 32), dtype=dtype, device=DEV)
        pb = torch.zeros_like(pa)
        pe = pe_c
    cu.attention(qkv, oa, S, N, heads, S // 2 if cross else 0, pe, pa)
    th.attention(qkv, ob, S, N, heads, S // 2 if cross else 0, pe, pb)
    rt, at = (1e-4, 2e-5) if dtype == torch.float32 else (6e-3, 4e-3)
    _close(oa, ob, rt, at, f"attention {case}")
    if pe_hw:
        _close(pa, pb, rt, at, f"pesum {case}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("C", [32, 128, 256, 384, 768])
def test_layernorm(ops, C, dtype):
    cu, th = ops
    g = torch.Generator().manual_seed(2)
    x = _rand((2, 3, 17, C + 8), dtype, g, 3.0)[..., :C]
    gamma, beta = _rand((C,), torch.float32, g), _rand((C,), torch.float32, g)
    for gb in ((None, None), (gamma, beta)):
        a = torch.zeros((2, 3, 17, C), dtype=dtype, device=DEV)
        b = torch.zeros_like(a)
        cu.layernorm(x, a, *gb)
        th.layernorm(x, b, *gb)
        _close(a, b, *((2e-5, 2e-5) if dtype == torch.float32 else (4e-3, 4e-3)), "layernorm")
    a32 = torch.zeros((2, 3, 17, C), dtype=torch.float32, device=DEV)
    b32 = torch.zeros_like(a32)
    cu.layernorm(x, a32, gamma, beta)
    th.layernorm(x, b32, gamma, beta)
    _close(a32, b32, 2e-5, 2e-5, "layernorm->fp32")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("C", [32, 128, 192])
def test_groupnorm(ops, C, dtype):
    cu, th = ops
    g = torch.Generator().manual_seed(3)
    x = _rand((2, 30, 44, C), dtype, g, 2.0) + 0.5
    gamma, beta = _rand((C,), torch.float32, g), _rand((C,), torch.float32, g)
    a = torch.zeros_like(x)
    b = torch.zeros_like(x)
    cu.groupnorm(x, 8, gamma, beta, a)
    th.groupnorm(x, 8, gamma, beta, b)
    _close(a, b, *((3e-5, 3e-5) if dtype == torch.float32 else (4e-3, 4e-3)), "groupnorm")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_resample_and_elementwise(ops, dtype):
    cu, th = ops
    g = torch.Generator().manual_seed(4)
    tol = (1e-6, 1e-6) if dtype == torch.float32 else (2e-3, 2e-3)
    x = _rand((2, 6, 10, 32 + 4), dtype, g)[..., 4:]
    for fn, shape in (("avgpool2", (2, 3, 5, 32)), ("upsample2x", (2, 12, 20, 32))):
        a = torch.zeros(shape, dtype=dtype, device=DEV)
        b = torch.zeros_like(a)
        getattr(cu, fn)(x, a)
        getattr(th, fn)(x, b)
        _close(a, b, *tol, fn)
    f, gg, z0, z1 = (_rand((2, 6, 10, 32), dtype, g) for _ in range(4))
    gg = torch.sigmoid(gg.float() * 3).to(dtype)
    a, b = torch.zeros_like(f), torch.zeros_like(f)
    cu.gate_blend(f, gg, z0, z1, a)
    th.gate_blend(f, gg, z0, z1, b)
    _close(a, b, *tol, "gate_blend")
    a, b = f.clone(), f.clone()
    cu.gate_blend(a, gg, z0, z1, a)      # in place on f, as the model does
    th.gate_blend(b, gg, z0, z1, b)
    _close(a, b, *tol, "gate_blend in place")
    for act in (ACT_TANH, ACT_GELU, ACT_SIGMOID, ACT_RELU):
        a, b = torch.zeros_like(f), torch.zeros_like(f)
        cu.unary(f, a, act)
        th.unary(f, b, act)
        _close(a, b, 2e-6 if dtype == torch.float32 else 2e-3, tol[1], f"unary {act}")
    a, b = torch.zeros_like(f), torch.zeros_like(f)
    cu.mul(f, z0, a)
    th.mul(f, z0, b)
    _close(a, b, *tol, "mul")
    ha, hb = z1.clone(), z1.clone()
    cu.gru_update(gg, f, ha)
    th.gru_update(gg, f, hb)
    _close(ha, hb, *tol, "gru_update")
    wide = torch.zeros((2, 6, 10, 72), dtype=dtype, device=DEV)
    cu.copy_channels(f, wide[..., 8:40])
    assert torch.equal(wide[..., 8:40], f) and float(wide[..., :8].abs().sum()) == 0 and float(wide[..., 40:].abs().sum()) == 0
    l, r = torch.randint(0, 256, (2, 3, 8, 12), generator=g).float().to(DEV), torch.randint(0, 256, (2, 3, 8, 12), generator=g).float().to(DEV)
    a = torch.zeros((4, 8, 12, 3), dtype=dtype, device=DEV)
    b = torch.zeros_like(a)
    cu.preprocess(l, r, a)
    th.preprocess(l, r, b)
    _close(a, b, *tol, "preprocess")


@pytest.mark.parametrize("pos", [True, False])
@pytest.mark.parametrize("w", [24, 40, 76, 161, 304, 608])   # 304 / 608: clusters of 4 / 8 CTAs (L, XL scan-lines)
def test_sinkhorn_softargmax(ops, w, pos):
    cu, th = ops
    g = torch.Generator().manual_seed(w)
    B, h = 2, 5
    cv = _rand((B, h, w, w), torch.float32, g, 6.0)
    outs = []
    for o in (cu, th):
        d, c, oc = (torch.zeros((B, h, w), device=DEV) for _ in range(3))
        am = torch.zeros((B, h, w), dtype=torch.int32, device=DEV)
        o.sinkhorn_softargmax(cv, d, c, oc, am, pos, 3)
        outs.append((d, c, oc, am))
    torch.cuda.synchronize()
    (d0, c0, o0, a0), (d1, c1, o1, a1) = outs
    assert torch.equal(a0, a1), f"argmax mismatches: {(a0 != a1).sum().item()}"
    _close(c0, c1, 1e-4, 1e-6, "conf")
    _close(o0, o1, 1e-4, 1e-6, "occ")
    _close(d0, d1, 1e-5, 2e-4, "disp")


@pytest.mark.parametrize("B,h,w,C", [(1, 3, 304, 256), (2, 2, 160, 128), (1, 2, 76, 192), (1, 1, 132, 64), (1, 2, 512, 256)])
def test_row_corr_split_fp16_tensor_core(ops, B, h, w, C):
    """LN(affine) + all-pairs scan-line correlation on tcgen05 with split-fp16 operands vs float64 torch: the error must stay at
    the level of an fp32 dot product (the arg-max over this volume is part of the bit-exact contract)."""
    cu, th = ops
    g = torch.Generator().manual_seed(B * 1000 + w + C)
    f = _rand((2 * B, h, w, C), torch.float16, g, 1.5) + 0.2
    gamma = (_rand((C,), torch.float32, g, 0.3) + 1.0).contiguous()
    beta = _rand((C,), torch.float32, g, 0.2).contiguous()
    cv = torch.full((B, h, w, w), 7.0, dtype=torch.float32, device=DEV)
    assert cu.row_corr_tc(f, gamma, beta, cv), "tcgen05 row correlation refused a supported shape"
    torch.cuda.synchronize()
    fn = torch.nn.functional.layer_norm(f.double(), (C,), gamma.double(), beta.double(), 1e-5)
    ref = torch.einsum("bhic,bhjc->bhij", fn[:B], fn[B:])
    err = (cv.double() - ref).abs().max().item()
    # the fp32 CUDA-core path (LayerNorm kernel -> batched fp32 GEMM) on the same input, as the yard-stick
    fn32 = torch.empty((2 * B, h, w, C), dtype=torch.float32, device=DEV)
    cu.layernorm(f, fn32, gamma, beta)
    cv32 = torch.empty_like(cv)
    cu.row_corr(fn32, cv32)
    err32 = (cv32.double() - ref).abs().max().item()
    print(f"row_corr split-fp16 B={B} h={h} w={w} C={C}: max err {err:.3e} (fp32 CUDA-core path: {err32:.3e}), |cv| max {ref.abs().max().item():.1f}")
    assert err < max(4 * err32, 2e-5 * ref.abs().max().item())


def test_cv_lookup_and_refiner_glue(ops):
    cu, th = ops
    g = torch.Generator().manual_seed(9)
    B, h, w = 2, 6, 37
    cv = _rand((B, h, w, w), torch.float32, g, 5.0)
    disp = (torch.rand((B, h, w), generator=g) * 45 - 4).to(DEV)
    a, b = torch.zeros((B, h, w, 18), device=DEV), torch.zeros((B, h, w, 18), device=DEV)
    cu.cv_lookup(cv, disp, a)
    th.cv_lookup(cv, disp, b)
    _close(a, b, 1e-5, 2e-5, "cv_lookup")
    conf = torch.rand((B, h, w), generator=g).to(DEV)
    occ = torch.rand((B, h, w), generator=g).to(DEV)
    for dtype in (torch.float32, torch.float16):
        oa = torch.zeros((B, h, w, 40), dtype=dtype, device=DEV)
        ob = torch.zeros_like(oa)
        cu.global_refiner_prep(disp, conf, oa[..., 32:34])
        th.global_refiner_prep(disp, conf, ob[..., 32:34])
        _close(oa, ob, 1e-5 if dtype == torch.float32 else 2e-3, 1e-6 if dtype == torch.float32 else 2e-3, "gr_prep")
    upd = _rand((B, h, w, 4), torch.float32, g, 10.0)
    da, db = disp.clone(), disp.clone()
    cu.global_refiner_merge(da, conf, upd, True)
    th.global_refiner_merge(db, conf, upd, True)
    _close(da, db, 0, 1e-6, "gr_merge")
    sa, sb = torch.zeros((B, h, w, 4), device=DEV), torch.zeros((B, h, w, 4), device=DEV)
    cu.local_refiner_prep(disp, conf, occ, sa)
    th.local_refiner_prep(disp, conf, occ, sb)
    _close(sa, sb, 2e-6, 2e-6, "lr_prep")
    upd3 = _rand((B, h, w, 8), torch.float32, g)
    xa = [t.clone() for t in (disp, conf, occ)]
    xb = [t.clone() for t in (disp, conf, occ)]
    cu.local_refiner_update(*xa, sa, upd3, True)
    th.local_refiner_update(*xb, sb, upd3, True)
    for u, v, nm in zip(xa, xb, ("disp", "conf", "occ")):
        _close(u, v, 2e-6, 2e-6, "lr_update " + nm)
    # the occlusion mask (s2m2.py:179) is a hard decision: identical zero pattern
    assert torch.equal(xa[2] == 0, xb[2] == 0)


def test_convex_upsampling(ops):
    cu, th = ops
    g = torch.Generator().manual_seed(10)
    B, h, w = 2, 5, 7
    disp, occ, conf = (torch.rand((B, h, w), generator=g).to(DEV) * s for s in (30, 1, 1))
    logits = _rand((B, 4 * h, 4 * w, 16), torch.float32, g, 2.0)     # 9 logits + 7 pad channels (ld = 16)
    ua, ub = torch.zeros((B, 4 * h, 4 * w, 4), device=DEV), torch.zeros((B, 4 * h, 4 * w, 4), device=DEV)
    cu.upsample4x_convex(disp, occ, conf, logits, ua)
    th.upsample4x_convex(disp, occ, conf, logits, ub)
    _close(ua, ub, 2e-6, 2e-5, "upsample4x")
    l1 = _rand((B, 4 * h, 4 * w, 9), torch.float32, g, 2.0)
    oa, ob = torch.zeros((3, B, 1, 4 * h, 4 * w), device=DEV), torch.zeros((3, B, 1, 4 * h, 4 * w), device=DEV)
    cu.filter1x(ub, l1, oa)
    th.filter1x(ub, l1, ob)
    _close(oa, ob, 2e-6, 2e-5, "filter1x")
    o2a, o2b = torch.zeros((3, B, 1, 8 * h, 8 * w), device=DEV), torch.zeros((3, B, 1, 8 * h, 8 * w), device=DEV)
    cu.filter1x(ub, l1, o2a, True)
    th.filter1x(ub, l1, o2b, True)
    _close(o2a, o2b, 2e-6, 4e-5, "filter1x output_upsample")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_fused_stem_and_mask1x_inputs(ops, dtype):
    cu, th = ops
    g = torch.Generator().manual_seed(12)
    B, H, W = 2, 10, 14
    l = torch.randint(0, 256, (B, 3, H, W), generator=g).float().to(DEV)
    r = torch.randint(0, 256, (B, 3, H, W), generator=g).float().to(DEV)
    cpu = lambda shape, sc: (torch.randn(shape, generator=g) * sc).contiguous()     # host tensors: kernel parameters
    w0, b0 = cpu((16, 3), 0.5), cpu((16,), 0.5)
    w1, b1 = cpu((16, 16), 0.25), cpu((16,), 0.25)
    a = torch.full((2 * B, H, W, 64), 5.0, dtype=dtype, device=DEV)       # channels 16..63 must come back as zeros
    b = torch.zeros_like(a)
    cu.stem(l, r, w0, b0, w1, b1, a)
    th.stem(l, r, w0, b0, w1, b1, b)
    tol = (2e-5, 2e-5) if dtype == torch.float32 else (3e-3, 3e-3)
    _close(a, b, *tol, "stem")
    up = _rand((B, H, W, 4), torch.float32, g, 20.0)
    rgb = _rand((B, H, W, 3), dtype, g)
    wd, bd = cpu((16, 9), 0.3), cpu((16,), 0.3)
    wr, br = cpu((16, 27), 0.2), cpu((16,), 0.2)
    oa = torch.zeros((B, H, W, 64), dtype=dtype, device=DEV)
    ob = torch.zeros_like(oa)
    cu.mask1x_inputs(up, rgb, wd, bd, wr, br, oa[..., 0:32])
    th.mask1x_inputs(up, rgb, wd, bd, wr, br, ob[..., 0:32])
    _close(oa, ob, *((2e-5, 2e-4) if dtype == torch.float32 else (3e-3, 3e-2)), "mask1x_inputs")


@pytest.mark.parametrize("C,M", [(256, 1000), (512, 300), (128, 2000), (384, 257), (192, 640), (256, 128 * 160)])
def test_fused_ffn(ops, C, M):
    """LN -> Linear -> GELU -> Linear -> +residual in one tcgen05 kernel vs the composite of torch ops (fp16 I/O)."""
    cu, th = ops
    g = torch.Generator().manual_seed(C + M)
    z = _rand((1, 1, M, C), torch.float16, g, 2.0) + 0.3
    l0 = ConvW(_rand((C, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((C,), torch.float32, g, 0.3), 1, 1, 1, 0, 0, C, C, False, "l0")
    l2 = ConvW(_rand((C, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((C,), torch.float32, g, 0.3), 1, 1, 1, 0, 0, C, C, False, "l2")
    ref = th.ffn_reference(z, l0, l2)
    za = z.clone()
    cu.ffn_min_tokens = cu.ffn_min_tokens_wide = 0
    assert cu.ffn(za, l0, l2), "fused FFN kernel refused a supported shape"
    torch.cuda.synchronize()
    _close(za, ref, 8e-3, 8e-3, f"fused ffn C={C} M={M}")


@pytest.mark.parametrize("C,M", [(256, 128 * 700 + 77), (256, 128 * 297), (128, 128 * 3), (192, 128 * 301 + 5), (256, 130)])
def test_fused_ffn_many_tiles(ops, C, M):
    """The two-tiles-in-flight FFN kernel over several tiles per CTA (odd / even counts, ragged tail), in place."""
    cu, th = ops
    g = torch.Generator().manual_seed(C * 7 + M)
    z = _rand((1, 1, M, C), torch.float16, g, 2.0) + 0.3
    l0 = ConvW(_rand((C, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((C,), torch.float32, g, 0.3), 1, 1, 1, 0, 0, C, C, False, "l0")
    l2 = ConvW(_rand((C, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((C,), torch.float32, g, 0.3), 1, 1, 1, 0, 0, C, C, False, "l2")
    ref = th.ffn_reference(z, l0, l2)
    za = z.clone()
    cu.ffn_min_tokens = 0
    assert cu.ffn(za, l0, l2)
    torch.cuda.synchronize()
    _close(za, ref, 8e-3, 8e-3, f"fused ffn (many tiles) C={C} M={M}")


@pytest.mark.parametrize("C,M,alias", [(256, 1000, False), (256, 128 * 600 + 3, True), (128, 2000, False), (192, 128 * 299, True)])
def test_fused_mlp2_convblock_branch(ops, C, M, alias):
    """1x1 -> ReLU -> 1x1 + res (ConvBlock2D's convs_1x branch summed with the 3x3 branch) in one kernel vs torch;
    strided views, output aliasing the residual or separate."""
    cu, th = ops
    g = torch.Generator().manual_seed(C * 3 + M)
    x = _rand((1, 1, M, C), torch.float16, g, 1.5)
    resb = _rand((1, 1, M, C + 8), torch.float16, g, 1.0)
    res = resb[..., :C]
    l0 = ConvW(_rand((C, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((C,), torch.float32, g, 0.3), 1, 1, 1, 0, 0, C, C, False, "x0")
    l2 = ConvW(_rand((C, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((C,), torch.float32, g, 0.3), 1, 1, 1, 0, 0, C, C, False, "x2")
    ref = th.mlp2_reference(x, res, l0, l2, ACT_RELU)
    guard = resb[..., C:].clone()
    out = res if alias else torch.full((1, 1, M, C), 9.0, dtype=torch.float16, device=DEV)
    cu.mlp2_min_tokens, cu.mlp2_max_tokens = 0, 1 << 30
    assert cu.mlp2(x, res, out, l0, l2, ACT_RELU), "fused mlp2 kernel refused a supported shape"
    torch.cuda.synchronize()
    _close(out, ref, 8e-3, 8e-3, f"mlp2 C={C} M={M}")
    assert torch.equal(resb[..., C:], guard), "mlp2 wrote outside its view"


@pytest.mark.parametrize("C,N,M,bias", [(256, 768, 1000, True), (128, 384, 2000, False), (192, 576, 640, True), (256, 768, 128 * 160, False),
                                        (128, 64, 129, True), (256, 320, 4096, True)])
def test_fused_ln_linear(ops, C, N, M, bias):
    """LayerNorm -> Linear (the QKV projection) in one tcgen05 kernel vs torch (fp16 I/O, fp32 accumulate)."""
    cu, th = ops
    g = torch.Generator().manual_seed(C + N + M)
    z = _rand((1, 1, M, C), torch.float16, g, 2.0) + 0.3
    lin = ConvW(_rand((N, C), torch.float16, g, 1.0 / math.sqrt(C)), _rand((N,), torch.float32, g, 0.3) if bias else None, 1, 1, 1, 0, 0, C, N,
                False, "qkv")
    ref = th.ln_linear_reference(z, lin)
    big = torch.full((1, 1, M, N + 8), 7.0, dtype=torch.float16, device=DEV)     # strided view: ld = N + 8, guard columns
    out = big[..., :N]
    cu.ln_linear_min_tokens = 0
    assert cu.ln_linear(z, lin, out), "fused LN+Linear kernel refused a supported shape"
    torch.cuda.synchronize()
    _close(out, ref, 6e-3, 6e-3, f"ln_linear C={C} N={N} M={M}")
    assert bool((big[..., N:] == 7.0).all()), "ln_linear wrote outside its view"


@pytest.mark.parametrize("C", [64, 192, 256, 512])
def test_resample_fp16_fast_paths(ops, C):
    """16-byte-per-thread avgpool2 / upsample2x kernels (aligned fp16 tensors, C % 8 == 0), including strided output views."""
    cu, th = ops
    g = torch.Generator().manual_seed(40 + C)
    x = _rand((3, 10, 14, C), torch.float16, g)
    for fn, shape in (("avgpool2", (3, 5, 7, C)), ("upsample2x", (3, 20, 28, C))):
        big = torch.full(shape[:3] + (C + 8,), 3.0, dtype=torch.float16, device=DEV)
        a, b = big[..., :C], torch.zeros(shape, dtype=torch.float16, device=DEV)
        getattr(cu, fn)(x, a)
        getattr(th, fn)(x, b)
        _close(a, b, 2e-3, 2e-3, f"{fn} C={C}")
        assert bool((big[..., C:] == 3.0).all()), f"{fn} wrote outside its view"


def test_stem_space_to_depth(ops):
    """fp16 stem output in space-to-depth layout (2B,H/2,W/2,64): channel (py*2+px)*16 + c of block (y/2, x/2)."""
    cu, th = ops
    g = torch.Generator().manual_seed(13)
    B, H, W = 2, 12, 20
    l = torch.randint(0, 256, (B, 3, H, W), generator=g).float().to(DEV)
    r = torch.randint(0, 256, (B, 3, H, W), generator=g).float().to(DEV)
    cpu = lambda shape, sc: (torch.randn(shape, generator=g) * sc).contiguous()
    w0, b0, w1, b1 = cpu((16, 3), 0.5), cpu((16,), 0.5), cpu((16, 16), 0.25), cpu((16,), 0.25)
    a = torch.full((2 * B, H // 2, W // 2, 64), 5.0, dtype=torch.float16, device=DEV)
    b = torch.zeros_like(a)
    cu.stem(l, r, w0, b0, w1, b1, a)
    th.stem(l, r, w0, b0, w1, b1, b)
    _close(a, b, 3e-3, 3e-3, "stem s2d")
    # and the layout itself against the pixel-major output
    c = torch.zeros((2 * B, H, W, 16), dtype=torch.float16, device=DEV)
    cu.stem(l, r, w0, b0, w1, b1, c)
    c = c.reshape(2 * B, H // 2, 2, W // 2, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(2 * B, H // 2, W // 2, 64)
    assert torch.equal(a, c)


def test_uint8_ingest_is_bit_identical(ops):
    """uint8 NCHW images through preprocess / stem give exactly the float-input results."""
    cu, _ = ops
    g = torch.Generator().manual_seed(21)
    B, H, W = 2, 16, 24
    lu = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).to(DEV)
    ru = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).to(DEV)
    cpu = lambda shape, sc: (torch.randn(shape, generator=g) * sc).contiguous()
    w0, b0, w1, b1 = cpu((16, 3), 0.5), cpu((16,), 0.5), cpu((16, 16), 0.25), cpu((16,), 0.25)
    for dtype in (torch.float16, torch.float32):
        a, b = (torch.zeros((2 * B, H, W, 3), dtype=dtype, device=DEV) for _ in range(2))
        cu.preprocess(lu, ru, a)
        cu.preprocess(lu.float(), ru.float(), b)
        assert torch.equal(a, b)
        a, b = (torch.zeros((2 * B, H, W, 16), dtype=dtype, device=DEV) for _ in range(2))
        cu.stem(lu, ru, w0, b0, w1, b1, a)
        cu.stem(lu.float(), ru.float(), w0, b0, w1, b1, b)
        assert torch.equal(a, b)


def test_post_confidence_score_and_points():
    """SURVEY §8(f)2,3: batched confidence score and disparity -> points vs the oracle restatement."""
    from oracle import post_oracle as PO
    from s2m2_b200 import post
    g = torch.Generator().manual_seed(5)
    conf = torch.rand((3, 1, 352, 416), generator=g)
    for hw, margin in (((352, 416), 100), ((301, 399), 100), ((340, 410), 7)):
        s = post.confidence_score(conf.to(DEV), hw, margin).cpu()
        assert torch.allclose(s, PO.confidence_score(conf, hw, margin), atol=1e-6, rtol=1e-6), (hw, margin)
    assert torch.equal(post.confidence_score(conf.to(DEV), (301, 399)), post.confidence_score(conf.to(DEV), (301, 399)))   # deterministic
    disp = torch.rand((2, 1, 40, 56), generator=g) * 60 - 3           # some non-positive disparities
    occ = torch.rand((2, 1, 40, 56), generator=g)
    cf = torch.rand((2, 1, 40, 56), generator=g)
    kw = dict(fx=711.5, cx=27.3, cy=19.1, baseline=193.0, doffs=31.2, depth_trunc=12.0)
    xyz, valid = post.disparity_to_points(disp.to(DEV), conf=cf.to(DEV), occ=occ.to(DEV), **kw)
    rx, rv = PO.disparity_to_points(disp, conf=cf, occ=occ, **kw)
    assert torch.equal(valid.cpu(), rv)
    assert torch.allclose(xyz.cpu(), rx, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("H,W,dtype", [(100, 75, torch.uint8), (64, 96, torch.float32), (70, 96, torch.uint8), (481, 641, torch.uint8)])
def test_image_pad_kernels(ops, H, W, dtype):
    """image_pad (image_utils.py:27-63) on the device vs its CPU restatement (which tests/test_post_cpu.py holds to the reference):
    the original pixels bit-identical, the blurred fill within fp32 summation-order noise."""
    from oracle import post_oracle as PO
    from s2m2_b200.pre import image_pad
    g = torch.Generator().manual_seed(H + W)
    img = torch.randint(0, 256, (2, 3, H, W), generator=g, dtype=torch.uint8).to(dtype)
    got = image_pad(img.to(DEV)).cpu()
    ref = PO.image_pad(img)
    assert got.shape == ref.shape
    hs, ws = (ref.shape[2] - H) // 2, (ref.shape[3] - W) // 2
    assert torch.equal(got[:, :, hs:hs + H, ws:ws + W], img.float())
    _close(got, ref, 1e-5, 2e-3, f"image_pad {H}x{W}")


def test_fused_crop_and_device_side_run_stereo_matching():
    """S2M2.forward(crop_hw=...) == image_crop of the full forward (bit-identical), and pre.stereo_match == the chain
    image_pad -> forward -> image_crop -> interior confidence mean of run_stereo_matching (model_utils.py:50-95)."""
    from oracle import post_oracle as PO
    from s2m2_b200.model import S2M2
    from s2m2_b200.pre import image_pad, stereo_match
    from s2m2_b200.weights import random_state_dict
    m = S2M2(32, 1, 1, use_positivity=True, refine_iter=2, precision="fp16").eval()
    m.my_load_state_dict(random_state_dict(32, 1, 0))
    m = m.to(DEV)
    g = torch.Generator().manual_seed(5)
    H, W = 70, 100
    left = torch.randint(0, 256, (2, 3, H, W), generator=g, dtype=torch.uint8).to(DEV)
    right = torch.randint(0, 256, (2, 3, H, W), generator=g, dtype=torch.uint8).to(DEV)
    lp, rp = image_pad(left), image_pad(right)
    full = m(lp, rp)
    crop = m(lp, rp, crop_hw=(H, W))
    for f, c in zip(full, crop):
        assert c.shape == (2, 1, H, W)
        assert torch.equal(PO.image_crop(f, (H, W)), c)
    d, o, c, score = stereo_match(m, left, right, margin=10)
    assert torch.equal(d, crop[0][:, 0]) and torch.equal(c, crop[2][:, 0])
    want = PO.confidence_score(full[2].cpu(), (H, W), 10)
    assert torch.allclose(score.cpu(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("emode,kh,kw,cin,cout,act", [(1, 3, 1, 128, 128, ACT_SIGMOID), (1, 1, 3, 256, 512, ACT_SIGMOID), (2, 1, 1, 256, 256, ACT_SIGMOID),
                                                     (2, 1, 1, 64, 64, ACT_SIGMOID), (3, 3, 1, 128, 64, ACT_TANH), (3, 1, 3, 512, 256, ACT_TANH)])
def test_conv_epilogue_modes(ops, emode, kh, kw, cin, cout, act):
    """The gate-blend / GRU epilogue modes of the tcgen05 conv (include/s2m2b200.h) vs the composition of separate torch ops;
    strided views, the GRU update written in place over R2."""
    cu, th = ops
    g = torch.Generator().manual_seed(emode * 100 + cin + cout)
    N, H, W = 2, 37, 53
    x = _rand((N, H, W, cin), torch.float16, g)
    w = _rand((cout, kh * kw * cin), torch.float16, g, 1.0 / math.sqrt(kh * kw * cin))
    cw = ConvW(w, _rand((cout,), torch.float32, g, 0.3), kh, kw, 1, kh // 2, kw // 2, cin, cout, False, f"emode{emode}")
    res_from = cout // 2 if emode == 1 else 0
    r1 = torch.rand((N, H, W, cout - res_from + 8), generator=g).to(DEV).half()[..., :cout - res_from]
    r2b = _rand((N, H, W, 2 * cout), torch.float16, g)
    r2 = r2b[..., :cout] if emode >= 2 else None
    ref = torch.empty((N, H, W, cout), dtype=torch.float16, device=DEV)
    th.conv2d(x, cw, act, ref, res=r1, res2=r2, emode=emode, res_from=res_from)
    out = r2 if emode == 3 else torch.full((N, H, W, cout), 9.0, dtype=torch.float16, device=DEV)    # mode 3: in place over R2
    guard = r2b[..., cout:].clone()
    cu.conv2d(x, cw, act, out, res=r1, res2=r2, emode=emode, res_from=res_from)
    torch.cuda.synchronize()
    _close(out, ref, 6e-3, 6e-3, f"conv epilogue mode {emode} {kh}x{kw} {cin}->{cout}")
    assert torch.equal(r2b[..., cout:], guard)


def test_remap_u8_bit_exact_with_opencv():
    """s2m2b200_remap_u8 (rectify_images on the device, image_utils.py:108-136) against cv2.remap's own outputs
    (tests/golden/remap_u8.npz, oracle/make_golden_remap.py): real rectification maps of two candidate rotations and a synthetic
    map with borders / integer coordinates / rounding ties — bit-exact, planar and interleaved, shared and per-candidate sources."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from s2m2_b200 import pre
    z = np.load(os.path.join(GOLDEN, "remap_u8.npz"))
    dev = torch.device("cuda", 0)
    left, right = torch.from_numpy(z["left"]).to(dev), torch.from_numpy(z["right"]).to(dev)
    cands = [{"leftMapX": z[f"c{i}_lx"], "leftMapY": z[f"c{i}_ly"], "rightMapX": z[f"c{i}_rx"], "rightMapY": z[f"c{i}_ry"]} for i in range(2)]
    l, r = pre.rectify_images(left, right, cands)                 # (2,3,H,W) planar: the forward's uint8 ingest layout
    li, ri = pre.rectify_images(left, right, cands, planar=False)
    for i in range(2):
        assert np.array_equal(l[i].permute(1, 2, 0).cpu().numpy(), z[f"c{i}_left"])
        assert np.array_equal(r[i].permute(1, 2, 0).cpu().numpy(), z[f"c{i}_right"])
        assert np.array_equal(li[i].cpu().numpy(), z[f"c{i}_left"]) and np.array_equal(ri[i].cpu().numpy(), z[f"c{i}_right"])
    one = pre.rectify_images(left, right, cands[1])[0]
    assert torch.equal(one[0], l[1])
    # synthetic map into a destination of another size; per-candidate sources through the op itself
    from s2m2_b200.ops import CudaOps
    ops = CudaOps()
    mx, my = torch.from_numpy(z["syn_mx"]).to(dev)[None], torch.from_numpy(z["syn_my"]).to(dev)[None]
    out = torch.empty((1,) + tuple(z["syn_out"].shape), dtype=torch.uint8, device=dev)
    ops.remap_u8(left, mx, my, out, planar=False)
    assert np.array_equal(out[0].cpu().numpy(), z["syn_out"])
    src2 = torch.stack([left, right])
    out2 = torch.empty((2,) + tuple(z["syn_out"].shape), dtype=torch.uint8, device=dev)
    ops.remap_u8(src2, mx.expand(2, -1, -1).contiguous(), my.expand(2, -1, -1).contiguous(), out2, planar=False)
    assert np.array_equal(out2[0].cpu().numpy(), z["syn_out"])
    from oracle.post_oracle import remap_u8
    assert np.array_equal(out2[1].cpu().numpy(), remap_u8(z["right"], z["syn_mx"], z["syn_my"]))


def test_candidate_scores_equal_one_by_one_evaluation():
    """pre.candidate_scores (N candidate rectifications -> N confidence scores, the calibration optimisers' inner loop,
    calibration/base.py:15-36, cem.py:64-72) equals rectifying and scoring each candidate on its own."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from s2m2_b200 import pre
    from s2m2_b200.model import S2M2
    from s2m2_b200.weights import random_state_dict
    z = np.load(os.path.join(GOLDEN, "remap_u8.npz"))
    dev = torch.device("cuda", 0)
    left, right = torch.from_numpy(z["left"]).to(dev), torch.from_numpy(z["right"]).to(dev)
    cands = [{"leftMapX": z[f"c{i}_lx"], "leftMapY": z[f"c{i}_ly"], "rightMapX": z[f"c{i}_rx"], "rightMapY": z[f"c{i}_ry"]} for i in range(2)]
    m = S2M2(32, 1, 1, use_positivity=True).eval()
    m.load_state_dict(random_state_dict(32, 1, 0))
    m = m.to(dev)
    scores = pre.candidate_scores(m, left, right, cands * 3, margin=20, chunk=4)
    assert scores.shape == (6,) and torch.isfinite(scores).all()
    for i in range(2):
        l, r = pre.rectify_images(left, right, cands[i])
        s = pre.stereo_match(m, l, r, margin=20)[3]
        assert torch.allclose(scores[i], s[0], atol=1e-6) and torch.allclose(scores[i + 2], s[0], atol=1e-6)
