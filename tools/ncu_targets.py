#!/usr/bin/env python
"""One launch of every round-2 kernel at its batch-8 L@1216x1024 shape (and the XL Sinkhorn of BASELINE config 4) — the target of
`ncu --set full`:

    ncu --set full --clock-control none --import-source on -k regex:"mlp2_tc|lnlinear_tc|attention2d_tc|attention_tc|sinkhorn_cluster|rowcorr_tc|ln_split" \
        -o gpurun_out/r2/prof_r2 python tools/ncu_targets.py
"""
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
    ops.ffn_min_tokens = ops.ln_linear_min_tokens = 0
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    M, C = 16 * 256 * 304, 256
    z = torch.randn((1, 1, M, C), generator=g).to(dev).half()
    mk = lambda n, k, bias=True: ConvW((torch.randn((n, k), generator=g) / math.sqrt(k)).to(dev).half(),
                                       torch.randn((n,), generator=g).to(dev) if bias else None, 1, 1, 1, 0, 0, k, n, False, "l")
    l0, l2, qw = mk(C, C), mk(C, C), mk(3 * C, C, False)
    qkv = torch.empty((1, 1, M, 3 * C), dtype=torch.float16, device=dev)
    for _ in range(2):
        assert ops.ffn(z, l0, l2)                                   # mlp2_tc_kernel<256, LN, GELU>: fused FFN, level 0, batch 8
        assert ops.ln_linear(z, qw, qkv)                            # lnlinear_tc_kernel<256>: LN + QKV
    S, N, heads, d = 16 * 256, 304, 1, 256                          # scan-line attention, level 0, batch 8
    o = torch.empty((1, S, N, C), dtype=torch.float16, device=dev)
    q4 = qkv.view(1, S, N, 3 * C)
    for _ in range(2):
        ops.attention(q4, o, S, N, heads, S // 2)
    del z, qkv, o, q4
    S, N, heads, d = 16, 1216, 8, 64                                # 2-D attention at 1/32, batch 8 (+ PE variant)
    qkv = torch.randn((1, S, N, 3 * heads * d), generator=g).to(dev).half()
    o = torch.empty((1, S, N, heads * d), dtype=torch.float16, device=dev)
    pe = ops.pe_tables(32, 38, torch.device(dev))
    pesum = torch.empty((S, N, heads, 32), dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.attention(qkv, o, S, N, heads, S // 2)
        ops.attention(qkv, o, S, N, heads, 0, pe, pesum)
    del qkv, o, pesum
    for name, B, h, w, Cc in (("L", 8, 256, 304, 256), ("XL", 1, 512, 608, 384)):    # DispInit: correlation + Sinkhorn
        f = (torch.randn((2 * B, h, w, Cc), generator=g) * 1.5).to(dev).half()
        gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
        cv = torch.empty((B, h, w, w), dtype=torch.float32, device=dev)
        if not ops.row_corr_tc(f, gamma, beta, cv):
            fn = torch.empty((2 * B, h, w, Cc), dtype=torch.float32, device=dev)
            ops.layernorm(f, fn, gamma, beta)
            ops.row_corr(fn, cv)
            del fn
        else:
            ops.row_corr_tc(f, gamma, beta, cv)
        dd, cc, oo = (torch.empty((B, h, w), device=dev) for _ in range(3))
        am = torch.empty((B, h, w), dtype=torch.int32, device=dev)
        for _ in range(2):
            ops.sinkhorn_softargmax(cv, dd, cc, oo, am, True, 3)
        del f, cv
    torch.cuda.synchronize()
    print("ncu targets done")


if __name__ == "__main__":
    main()
