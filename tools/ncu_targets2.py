#!/usr/bin/env python
"""One launch (after a warm-up launch) of the kernels round 1's review found without ncu evidence, at their batch-8 L@1216x1024
shapes: the K <= 512 1x1 family of the tcgen05 conv (`conv_tc_persistent_kernel<256,3,0>`), its 64-wide variant (`<64,8,0>`),
`conv_smallk`, and the HBM-bound glue (GroupNorm, avg-pool, bilinear x2, cost-volume lookup, convex 4x upsampling, the 1x filter).

    ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persistent|conv_smallk|gn_|avgpool2|upsample2x|cv_lookup|upsample4x|filter1x" \
        -o gpurun_out/r2/prof_r2c python tools/ncu_targets2.py
    python tools/ncu_summary.py gpurun_out/r2/prof_r2c.ncu-rep
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import ACT_GELU, ACT_NONE, ACT_RELU, ConvW  # noqa: E402
from s2m2_b200.ops import CudaOps  # noqa: E402


def main():
    ops = CudaOps()
    g = torch.Generator().manual_seed(0)
    dev = torch.device("cuda", 0)
    f16, f32 = torch.float16, torch.float32

    def mk(n, k, kh=1, kw=1, bias=True):
        K = kh * kw * k
        return ConvW((torch.randn((n, K), generator=g) / math.sqrt(K)).to(dev).half(), torch.randn((n,), generator=g).to(dev) if bias else None,
                     kh, kw, 1, kh // 2, kw // 2, k, n, False, f"k{kh}x{kw}_{k}_{n}")

    B2, h4, w4, C = 16, 256, 304, 256                                  # both images of 8 pairs at 1/4 resolution
    z0 = torch.randn((B2, h4, w4, C), generator=g).to(dev).half()
    z1 = torch.randn((B2, h4, w4, C), generator=g).to(dev).half()
    t = torch.empty((B2, h4, w4, 3 * C), dtype=f16, device=dev)
    o = torch.empty((B2, h4, w4, C), dtype=f16, device=dev)
    gf0, f2, proj = mk(3 * C, 2 * C), mk(C, 2 * C), mk(C, C, bias=False)
    for _ in range(2):
        ops.conv2d(z0, gf0, ACT_GELU, t, None, z1)                     # FeatureFusion gate|fusion hidden: 1x1, K = 512 (two sources), N = 768, GELU
        ops.conv2d(t[..., C:], f2, ACT_NONE, o, z0)                    # feature_fusion.2: 1x1, K = 512 (strided view), N = 256, + residual
        ops.conv2d(z1, proj, ACT_NONE, z0, z0)                         # attention proj: 1x1, K = 256, N = 256, + residual, in place
    del t, o, z1
    # GroupNorm(8) over the 1/2-resolution backbone map; avg-pool / bilinear x2 of the U-Nets
    x = torch.randn((B2, 512, 608, C), generator=g).to(dev).half()
    y = torch.empty_like(x)
    gw, gb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    p = torch.empty((B2, h4 // 2, w4 // 2, C), dtype=f16, device=dev)
    u = torch.empty((B2, 2 * h4, 2 * w4, C), dtype=f16, device=dev)
    for _ in range(2):
        ops.groupnorm(x, 8, gw, gb, y)
        ops.avgpool2(z0, p)
        ops.upsample2x(z0, u)
    del x, y, p, u, z0
    # 64-wide layers at full resolution (mask head), conv_smallk (1 -> 128 channels from the fp32 disparity map)
    B, H, W = 8, 1024, 1216
    m1 = torch.randn((B, H, W, 64), generator=g).to(dev).half()
    m2 = torch.empty_like(m1)
    c0 = mk(64, 64, 3, 3)
    small = torch.randn((B, h4, w4, 4), generator=g).to(dev)
    df0 = mk(128, 1, 3, 3)
    d0 = torch.empty((B, h4, w4, 128), dtype=f16, device=dev)
    for _ in range(2):
        ops.conv2d(m1, c0, ACT_RELU, m2)                               # upsample_mask_1x.conv_concat.0: 3x3, 64 -> 64, full resolution
        ops.conv2d(small[..., 0:1], df0, ACT_GELU, d0)                 # refiner.disp_feat.0: 3x3, 1 -> 128 (conv_smallk)
    del m1, m2
    # matching glue: cost-volume lookup, convex 4x upsampling, edge-guided 1x filter
    cv = torch.randn((B, h4, w4, w4), generator=g).to(dev)
    disp = (torch.rand((B, h4, w4), generator=g) * 60).to(dev)
    conf, occ = torch.rand((B, h4, w4), generator=g).to(dev), torch.rand((B, h4, w4), generator=g).to(dev)
    corr = torch.empty((B, h4, w4, 18), dtype=f32, device=dev)
    lg4 = torch.randn((B, H, W, 16), generator=g).to(dev)
    up = torch.empty((B, H, W, 4), dtype=f32, device=dev)
    out = torch.empty((3, B, 1, H, W), dtype=f32, device=dev)
    for _ in range(2):
        ops.cv_lookup(cv, disp, corr)
        ops.upsample4x_convex(disp, occ, conf, lg4, up)
        ops.filter1x(up, lg4, out)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
