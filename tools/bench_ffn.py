#!/usr/bin/env python
"""Micro-benchmark: fused LN+FFN tcgen05 kernel vs LayerNorm + two conv2d launches, at the L@1216x1024 level shapes."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import ACT_GELU, ACT_NONE, ConvW  # noqa: E402
from s2m2_b200.ops import CudaOps  # noqa: E402

SHAPES = [("level0 1/4", 155648, 256), ("level1 1/8", 38912, 256), ("level2 1/16", 9728, 512), ("level3 1/32", 2432, 512),
          ("level0 b8", 1245184, 256), ("level1 b8", 311296, 256), ("level2 b8", 77824, 512), ("level3 b8", 19456, 512)]


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
    for name, M, C in SHAPES:
        z = (torch.randn((1, 1, M, C), generator=g)).cuda().half()
        mk = lambda: ConvW((torch.randn((C, C), generator=g) / math.sqrt(C)).cuda().half(), torch.randn((C,), generator=g).cuda(),
                           1, 1, 1, 0, 0, C, C, False, "l")
        l0, l2 = mk(), mk()
        zn, h = torch.empty_like(z), torch.empty_like(z)

        def unfused():
            ops.layernorm(z, zn)
            ops.conv2d(zn, l0, ACT_GELU, h)
            ops.conv2d(h, l2, ACT_NONE, z, res=z)

        ops.ffn_min_tokens = ops.ffn_min_tokens_wide = ops.mlp2_min_tokens = 0
        ops.mlp2_max_tokens = 1 << 30
        t_f = timeit(lambda: ops.ffn(z, l0, l2))
        t_u = timeit(unfused)
        gf = 4.0 * M * C * C / 1e9
        print(f"{name:14s} M={M:7d} C={C:4d}   fused {t_f:8.1f} us ({gf / t_f * 1e-3:6.0f} TFLOP/s, {4.0 * M * C / t_f * 1e-3:5.0f} GB/s)   "
              f"LN+2 GEMMs {t_u:8.1f} us", flush=True)
        if C <= 256:
            a = torch.empty_like(z)
            from s2m2_b200.model import ACT_RELU

            def unfused_cb():
                ops.conv2d(z, l0, ACT_RELU, h)
                ops.conv2d(h, l2, ACT_NONE, a, res=a)
            t_f = timeit(lambda: ops.mlp2(z, a, a, l0, l2, ACT_RELU))
            t_u = timeit(unfused_cb)
            print(f"{'':14s} convblock 1x1 branch    fused {t_f:8.1f} us   2 GEMMs    {t_u:8.1f} us", flush=True)
        if C <= 256:
            qw = ConvW((torch.randn((3 * C, C), generator=g) / math.sqrt(C)).cuda().half(), None, 1, 1, 1, 0, 0, C, 3 * C, False, "qkv")
            qkv = torch.empty((1, 1, M, 3 * C), dtype=torch.float16, device="cuda")
            ops.ln_linear_min_tokens = 0

            def unfused_q():
                ops.layernorm(z, zn)
                ops.conv2d(zn, qw, ACT_NONE, qkv)

            t_f = timeit(lambda: ops.ln_linear(z, qw, qkv))
            t_u = timeit(unfused_q)
            print(f"{'':14s} LN+QKV  N={3 * C:4d}   fused {t_f:8.1f} us   LN+GEMM    {t_u:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
