#!/usr/bin/env python
"""Micro-benchmark of the tiny-K encoder convolutions of the local refiner (refinenet.py:87-100) at 1/4 resolution of L@1216x1024."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import ACT_GELU, ConvW  # noqa: E402
from s2m2_b200.ops import CudaOps  # noqa: E402


def main():
    ops = CudaOps()
    g = torch.Generator().manual_seed(0)
    for B in (1, 8):
        for name, cin, k, ld, dt in (("disp 1->128 3x3 (fp32 map, ld 4)", 1, 3, 4, torch.float32), ("conf/occ 2->128 3x3 (fp32, ld 4)", 2, 3, 4, torch.float32),
                                     ("corr 9->128 1x1 (fp16, ld 18)", 9, 1, 18, torch.float16)):
            K = k * k * cin
            w = (torch.randn((128, K), generator=g) / math.sqrt(K)).cuda().half()
            b = torch.randn((128,), generator=g).cuda()
            cw = ConvW(w, b, k, k, 1, k // 2, k // 2, cin, 128, False, name)
            xb = torch.randn((B, 256, 304, ld), generator=g).cuda().to(dt)
            x = xb[..., :cin]
            y = torch.empty((B, 256, 304, 128), dtype=torch.float16, device="cuda")
            for _ in range(3):
                ops.conv2d(x, cw, ACT_GELU, y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                ops.conv2d(x, cw, ACT_GELU, y)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            print(f"batch {B} {name:36s} {us:8.1f} us   out {y.numel() * 2 / us / 1e3:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
