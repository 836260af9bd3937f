#!/usr/bin/env python
"""Device time of the small HBM-bound glue kernels at their batch-8 L@1216x1024 shapes (CUDA events, 20 launches after 3)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import ACT_GELU, ConvW  # noqa: E402
from s2m2_b200.ops import CudaOps  # noqa: E402


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
    dev = torch.device("cuda", 0)
    for shape in ((16, 128, 152, 256), (16, 64, 76, 512), (16, 256, 304, 256)):
        x = torch.randn(shape, generator=g).to(dev).half()
        u = torch.empty((shape[0], 2 * shape[1], 2 * shape[2], shape[3]), dtype=torch.float16, device=dev)
        t = timeit(lambda: ops.upsample2x(x, u))
        gb = (x.numel() + u.numel()) * 2 / 1e9
        print(f"upsample2x {shape}: {t:8.1f} us  {gb / t * 1e6:7.0f} GB/s", flush=True)
        del x, u
    B, h4, w4 = 8, 256, 304
    small = torch.randn((B, h4, w4, 4), generator=g).to(dev)
    corr = torch.randn((B, h4, w4, 18), generator=g).to(dev)

    def mk(n, k, kh, kw):
        K = kh * kw * k
        return ConvW((torch.randn((n, K), generator=g) / math.sqrt(K)).to(dev).half(), torch.randn((n,), generator=g).to(dev), kh, kw, 1, kh // 2, kw // 2, k, n, False, "s")
    for name, x, cw in (("disp_feat.0 3x3 1->128", small[..., 0:1], mk(128, 1, 3, 3)), ("corr_feat1.0 1x1 9->128", corr[..., 0:9], mk(128, 9, 1, 1)),
                        ("conf_occ_feat.0 3x3 2->64", small[..., 1:3], mk(64, 2, 3, 3))):
        o = torch.empty((B, h4, w4, cw.cout), dtype=torch.float16, device=dev)
        t = timeit(lambda: ops.conv2d(x, cw, ACT_GELU, o))
        print(f"conv_smallk {name}: {t:8.1f} us  (output {o.numel() * 2 / 1e6:.0f} MB -> {o.numel() * 2 / t / 1e3:.0f} GB/s)", flush=True)


def narrow_convs():
    """The 64-wide tcgen05 layers (S2M2_TC_BRES=0 streams their weights through the ring, the default keeps them resident)."""
    from s2m2_b200.model import ACT_NONE, ACT_RELU
    ops = CudaOps()
    g = torch.Generator().manual_seed(0)
    dev = torch.device("cuda", 0)

    def mk(n, k, kh, kw):
        K = kh * kw * k
        return ConvW((torch.randn((n, K), generator=g) / math.sqrt(K)).to(dev).half(), torch.randn((n,), generator=g).to(dev), kh, kw, 1, kh // 2, kw // 2, k, n, False, "n")
    cases = (("upsample_mask_1x.conv_concat.0 3x3 64->64 @1x, 8 img", (8, 1024, 1216, 64), mk(64, 64, 3, 3), ACT_RELU, torch.float16),
             ("upsample_mask_1x.conv_concat.2 1x1 64->16 fp32 @1x", (8, 1024, 1216, 64), mk(16, 64, 1, 1), ACT_NONE, torch.float32),
             ("cnn_backbone.conv1_down.0 3x3 64->64 @1/2, 16 img", (16, 512, 608, 64), mk(64, 64, 3, 3), ACT_GELU, torch.float16))
    for name, shape, cw, act, odt in cases:
        x = torch.randn(shape, generator=g).to(dev).half()
        o = torch.empty(shape[:3] + (cw.cout,), dtype=odt, device=dev)
        t = timeit(lambda: ops.conv2d(x, cw, act, o))
        gb = (x.numel() * 2 + o.numel() * o.element_size()) / 1e9
        fl = 2.0 * x.numel() / shape[3] * cw.cout * cw.kh * cw.kw * cw.cin
        print(f"BRES={os.environ.get('S2M2_TC_BRES', '1')} {name}: {t:8.1f} us  {gb / t * 1e6:6.0f} GB/s  {fl / t / 1e6:6.0f} TFLOP/s", flush=True)
        del x, o


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "narrow":
        narrow_convs()
        sys.exit(0)
    main()
