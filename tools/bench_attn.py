#!/usr/bin/env python
"""Micro-benchmark of the attention kernels at the L@1216x1024 level shapes (S sequences, N tokens, heads, head dim)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.ops import CudaOps  # noqa: E402

#          name                 S    N    heads d    cross
SHAPES = [("1/4  scan-line",    512, 304, 1,    256, True),
          ("1/8  scan-line",    256, 152, 2,    128, True),
          ("1/16 scan-line",    128, 76,  4,    128, True),
          ("1/32 global 2-D",   2,   1216, 8,   64,  True),
          ("1/32 2-D batch 2",  4,   1216, 8,   64,  True),
          ("1/32 2-D batch 4",  8,   1216, 8,   64,  True),
          ("1/32 2-D batch 8",  16,  1216, 8,   64,  True),
          ("1/4  b8 scan-line", 4096, 304, 1,   256, True),
          ("1/8  b8 scan-line", 2048, 152, 2,   128, True),
          ("1/32 2-D b8 PE",    16,  1216, 8,   64,  False)]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    ops = CudaOps()
    g = torch.Generator().manual_seed(0)
    for name, S, N, heads, d, cross in SHAPES:
        C = heads * d
        qkv = torch.randn((1, S, N, 3 * C), generator=g).cuda().half()
        o = torch.empty((1, S, N, C), dtype=torch.float16, device="cuda")
        pe = pesum = None
        if name.endswith("PE"):
            pe = ops.pe_tables(32, 38, torch.device("cuda"))
            pesum = torch.empty((S, N, heads, 32), dtype=torch.float16, device="cuda")
        t = timeit(lambda: ops.attention(qkv, o, S, N, heads, S // 2 if cross else 0, pe, pesum))
        flops = 4.0 * S * heads * N * N * d
        byts = qkv.numel() * 2 + o.numel() * 2
        print(f"{name:18s} S={S:4d} N={N:5d} h={heads} d={d:3d}  {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
