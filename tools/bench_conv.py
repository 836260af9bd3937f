#!/usr/bin/env python
"""Micro-benchmark of the implicit-GEMM conv kernels on representative S2M2 layer shapes (L@1216x1024).

    python tools/bench_conv.py [--impl 0|1|2] [--reps 20] [--only NAME]
Prints one line per shape: time (CUDA events) and TFLOP/s.  Run under `ncu -k regex:conv_tc ...` for captures.
"""
import argparse
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import ACT_GELU, ACT_NONE, ConvW  # noqa: E402
from s2m2_b200.ops import CudaOps  # noqa: E402

SHAPES = [
    # name, N, H, W, Cin, Cout, k, stride, act, residual
    ("bb.conv2 3x3 1/2res C256", 2, 512, 608, 256, 256, 3, 1, ACT_GELU, False),
    ("bb.conv2 3x3 1/2res C256 batch8", 16, 512, 608, 256, 256, 3, 1, ACT_GELU, False),
    ("convblock 3x3 1/4res C256", 2, 256, 304, 256, 256, 3, 1, ACT_GELU, False),
    ("ffn 1x1 1/4res C256", 2, 256, 304, 256, 256, 1, 1, ACT_GELU, False),
    ("qkv 1x1 1/4res 256->768", 2, 256, 304, 256, 768, 1, 1, ACT_NONE, False),
    ("proj 1x1 1/4res +res", 2, 256, 304, 256, 256, 1, 1, ACT_NONE, True),
    ("convblock 3x3 1/8res C256", 2, 128, 152, 256, 256, 3, 1, ACT_GELU, False),
    ("ffn 1x1 1/8res C256", 2, 128, 152, 256, 256, 1, 1, ACT_GELU, False),
    ("convblock 3x3 1/16res C512", 2, 64, 76, 512, 512, 3, 1, ACT_GELU, False),
    ("ffn 1x1 1/16res C512", 2, 64, 76, 512, 512, 1, 1, ACT_GELU, False),
    ("ffn 1x1 1/32res C512", 2, 32, 38, 512, 512, 1, 1, ACT_GELU, False),
    ("gru 3x1 1/4res 512->512", 1, 256, 304, 512, 512, 3, 1, ACT_NONE, False),
    ("down 3x3 s2 1/2->1/4 C256", 2, 512, 608, 256, 256, 3, 2, ACT_NONE, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    ops = CudaOps(conv_impl=a.impl)
    g = torch.Generator().manual_seed(0)
    for name, N, H, W, Cin, Cout, k, stride, act, use_res in SHAPES:
        if a.only and a.only not in name:
            continue
        kh, kw = (3, 1) if "3x1" in name else (k, k)
        K = kh * kw * Cin
        w = (torch.randn((Cout, K), generator=g) / math.sqrt(K)).cuda().half()
        b = torch.randn((Cout,), generator=g).cuda()
        cw = ConvW(w, b, kh, kw, stride, kh // 2, kw // 2, Cin, Cout, False, name)
        x = torch.randn((N, H, W, Cin), generator=g).cuda().half()
        Ho, Wo = H // stride, W // stride
        y = torch.empty((N, Ho, Wo, Cout), dtype=torch.float16, device="cuda")
        res = torch.randn((N, Ho, Wo, Cout), generator=g).cuda().half() if use_res else None
        for _ in range(3):
            ops.conv2d(x, cw, act, y, res)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            ops.conv2d(x, cw, act, y, res)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        flops = 2.0 * N * Ho * Wo * Cout * K
        print(f"{name:32s} M={N * Ho * Wo:7d} N={Cout:4d} K={K:5d}  {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
