#!/usr/bin/env python
"""Micro-benchmark of the DispInit stage (submodules.py:203-243) at the BASELINE shapes: LayerNorm + split-fp16 tcgen05
correlation, then the cluster-resident Sinkhorn / soft-argmax, against the round-1 kernels (fp32 CUDA-core GEMM, streaming Sinkhorn).

    python tools/bench_dispinit.py [--cs 1|2|4|8] [--model L|S|XL] [--batch 8]
"""
import argparse
import os
import sys

ap = argparse.ArgumentParser()
ap.add_argument("--cs", type=int, default=0, help="force the Sinkhorn cluster size (S2M2_SK_CS)")
ap.add_argument("--model", default="L")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
if args.cs:
    os.environ["S2M2_SK_CS"] = str(args.cs)

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.ops import CudaOps  # noqa: E402

SHAPES = {"S": (128, 120, 160), "L": (256, 256, 304), "XL": (384, 512, 608)}


def timeit(fn, reps):
    for _ in range(2):
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
    C, h, w = SHAPES[args.model]
    B = args.batch
    g = torch.Generator().manual_seed(0)
    f = (torch.randn((2 * B, h, w, C), generator=g) * 1.5).cuda().half()
    gamma = torch.ones(C, device="cuda")
    beta = torch.zeros(C, device="cuda")
    cv = torch.empty((B, h, w, w), dtype=torch.float32, device="cuda")
    d, c, o = (torch.empty((B, h, w), device="cuda") for _ in range(3))
    am = torch.empty((B, h, w), dtype=torch.int32, device="cuda")
    cv_bytes = cv.numel() * 4
    if ops.row_corr_tc(f, gamma, beta, cv):
        t = timeit(lambda: ops.row_corr_tc(f, gamma, beta, cv), args.reps)
        print(f"{args.model} b{B}: LN + split-fp16 tcgen05 correlation   {t:8.1f} us   (cv write {cv_bytes / t / 1e3:7.0f} GB/s)", flush=True)
    fn = torch.empty((2 * B, h, w, C), dtype=torch.float32, device="cuda")

    def old():
        ops.layernorm(f, fn, gamma, beta)
        ops.row_corr(fn, cv)
    t = timeit(old, max(2, args.reps // 3))
    print(f"{args.model} b{B}: LayerNorm + fp32 CUDA-core correlation   {t:8.1f} us", flush=True)
    t = timeit(lambda: ops.sinkhorn_softargmax(cv, d, c, o, am, True, 3), args.reps)
    print(f"{args.model} b{B}: Sinkhorn (S2M2_SK_IMPL={os.environ.get('S2M2_SK_IMPL', '0')}, CS={args.cs or 'auto'})   {t:8.1f} us   "
          f"(cv read once: {cv_bytes / t / 1e3:7.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
