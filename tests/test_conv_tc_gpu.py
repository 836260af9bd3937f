"""The tcgen05/TMA implicit-GEMM convolution (forced, impl=2) vs the plain-torch op, fp16 operands."""
import math

import pytest
import torch

from s2m2_b200.model import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ConvW
from torch_ops import TorchOps

pytestmark = pytest.mark.gpu
DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

CASES = [
    # N, H, W, C1, C2, Cout, kh, kw, act, res, shuffle, out_f32, in_ld_extra
    (1, 4, 300, 256, 0, 256, 1, 1, ACT_GELU, False, False, False, 0),     # linear, flat tiles
    (2, 16, 38, 128, 0, 128, 1, 1, ACT_NONE, True, False, False, 0),      # + residual, ragged M
    (1, 24, 40, 64, 0, 64, 3, 3, ACT_GELU, False, False, False, 0),       # 3x3, BN=64
    (2, 19, 37, 128, 0, 256, 3, 3, ACT_RELU, True, False, False, 64),     # 3x3 ragged patch, strided input view
    (1, 32, 48, 64, 64, 192, 1, 1, ACT_GELU, False, False, False, 0),     # two-source 1x1, Cout 192 (BN=128, 2 tiles)
    (1, 20, 28, 128, 128, 384, 3, 3, ACT_SIGMOID, False, False, False, 0),  # two-source 3x3 (feat_fusion_layer)
    (1, 24, 30, 256, 0, 128, 3, 1, ACT_TANH, False, False, False, 0),     # GRU 3x1
    (1, 24, 30, 256, 0, 256, 1, 3, ACT_SIGMOID, False, False, False, 0),  # GRU 1x3, N=256
    (1, 16, 20, 128, 0, 4 * 64, 1, 1, ACT_NONE, False, True, False, 0),   # convT 2x2 s2 (Cr=64)
    (1, 16, 20, 128, 0, 4 * 16, 1, 1, ACT_NONE, False, True, False, 0),   # convT 2x2 s2 (Cr=16)
    (1, 16, 24, 128, 0, 96, 3, 3, ACT_NONE, False, False, True, 0),       # fp32 output, Cout=96
    (1, 8, 152, 512, 0, 512, 1, 1, ACT_GELU, True, False, False, 0),      # K=512, N=512
    (2, 32, 38, 512, 0, 512, 3, 3, ACT_NONE, False, False, False, 0),     # 1/32-res style 3x3, K=4608
    (2, 24, 36, 128, 0, 128, 3, 3, ACT_NONE, False, False, False, 0, 2),  # 3x3 stride 2 (conv2_down) via parity views
    (1, 64, 76, 256, 0, 256, 3, 3, ACT_GELU, False, False, False, 0, 2),  # 3x3 stride 2, N=256
    (2, 48, 72, 64, 0, 64, 5, 5, ACT_GELU, False, False, False, 0, 2),    # 5x5 stride 2 (conv1_down.0 with padded channels)
    # CTA-pair (cta_group::2) path: BN = 256, >= 296 tiles, K >= 1024
    (2, 152, 127, 64, 0, 256, 3, 3, ACT_GELU, True, False, False, 0),     # 3x3 + residual, ragged patches, odd number of m-tiles
    (1, 1, 39001, 1024, 0, 768, 1, 1, ACT_NONE, False, False, False, 0),  # flat 1x1, ragged M, 3 n-tiles, odd number of m-tiles
    (1, 160, 250, 512, 512, 384, 1, 1, ACT_GELU, False, False, False, 0), # two-source 1x1, Cout 384 (second n-tile half empty)
    (2, 300, 260, 64, 0, 256, 3, 3, ACT_RELU, False, False, False, 0, 2), # 3x3 stride 2 through the parity views
    (1, 200, 200, 128, 0, 256, 3, 1, ACT_TANH, True, False, True, 0),     # GRU 3x1, fp32 output (no TMA store), residual loads
]


@pytest.mark.parametrize("case", CASES)
def test_conv_tc(case):
    from s2m2_b200.ops import CudaOps
    cu, th = CudaOps(conv_impl=2), TorchOps()
    N, H, W, C1, C2, Cout, kh, kw, act, use_res, shuffle, out_f32, ldx = case[:13]
    stride = case[13] if len(case) > 13 else 1
    g = torch.Generator().manual_seed(abs(hash(case)) % 1000)
    Cin = C1 + C2
    K = kh * kw * Cin
    w = (torch.randn((Cout, K), generator=g) / math.sqrt(K)).to(DEV).half()
    b = (torch.randn((Cout,), generator=g) * 0.5).to(DEV)
    cw = ConvW(w, b, kh, kw, stride, kh // 2, kw // 2, Cin, Cout, shuffle, "tc")
    xb = torch.randn((N, H, W, C1 + ldx), generator=g).to(DEV).half()
    x = xb[..., :C1] if ldx else xb
    x2 = torch.randn((N, H, W, C2), generator=g).to(DEV).half() if C2 else None
    oshape = (N, 2 * H, 2 * W, Cout // 4) if shuffle else (N, H // stride, W // stride, Cout)
    odt = torch.float32 if out_f32 else torch.float16
    res = torch.randn(oshape, generator=g).to(DEV).half() if use_res else None
    a = torch.full(oshape, 7.0, dtype=odt, device=DEV)
    r = torch.zeros(oshape, dtype=odt, device=DEV)
    cu.conv2d(x, cw, act, a, res, x2)
    torch.cuda.synchronize()
    th.conv2d(x, cw, act, r, res, x2)
    err = (a.float() - r.float()).abs()
    bound = 8e-3 + 6e-3 * r.float().abs()
    assert torch.isfinite(a).all()
    assert bool((err <= bound).all()), f"{case}: max err {err.max().item():.3e}, frac bad {(err > bound).float().mean().item():.3e}"
