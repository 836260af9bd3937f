#!/usr/bin/env python
"""Per-layer device time of one forward (eager, CUDA events around every op call), grouped by layer name and shape.

    python tools/profile_layers.py [--model L --height 1024 --width 1216 --precision fp16] [-o profiles/rN_layers.md]

Complements the ncu launch list (which names kernels, not layers): this names the reference module each microsecond
belongs to.  Event pairs add ~2 us per op, so use it for ranking, not for absolute totals.
"""
import argparse
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import S2M2  # noqa: E402
from s2m2_b200.ops import CudaOps  # noqa: E402
from s2m2_b200.weights import MODEL_TABLE, random_state_dict, synthetic_pair  # noqa: E402

TIMED = ["conv2d", "ffn", "ln_linear", "row_corr", "attention", "layernorm", "groupnorm", "avgpool2", "upsample2x", "copy_channels", "unary",
         "gate_blend", "mul", "gru_update", "preprocess", "stem", "mask1x_inputs", "sinkhorn_softargmax", "cv_lookup", "global_refiner_prep",
         "global_refiner_merge", "local_refiner_prep", "local_refiner_update", "upsample4x_convex", "filter1x"]


class TimedOps(CudaOps):
    def __init__(self):
        super().__init__()
        self.records = []
        self.on = False
        for name in TIMED:
            setattr(self, name, self._wrap(name, getattr(super(), name)))

    def _wrap(self, name, fn):
        def call(*a, **k):
            if not self.on:
                return fn(*a, **k)
            label = name
            for v in a:
                if hasattr(v, "kh") and hasattr(v, "name"):
                    label = f"{name}:{v.name}"
                    break
            t = next((v for v in a if isinstance(v, torch.Tensor)), None)
            shape = tuple(t.shape) if t is not None else ()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            if r is not False:
                self.records.append((label, shape, e0, e1))
            return r
        return call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="L")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1216)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("-o", "--out", default=None)
    a = ap.parse_args()
    C, NTR = MODEL_TABLE[a.model]
    m = S2M2(C, 1, NTR, use_positivity=True, refine_iter=3, precision=a.precision).eval()
    m.my_load_state_dict(random_state_dict(C, NTR, 0))
    m = m.cuda()
    m.ops = TimedOps()
    m.auto_graph = False      # per-op CUDA events need eager launches
    left, right = synthetic_pair(a.batch, a.height, a.width, 1)
    left, right = left.cuda(), right.cuda()
    with torch.no_grad():
        for _ in range(3):
            m(left, right)
        torch.cuda.synchronize()
        m.ops.on = True
        m(left, right)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for label, shape, e0, e1 in m.ops.records:
        # strip the block index so the NTR x blocks of one kind add up:  mrt.2.blocks.1.self.qkv -> mrt.*.blocks.*.self.qkv
        key = (".".join("*" if s.isdigit() else s for s in label.split(".")), shape)
        agg[key][0] += 1
        agg[key][1] += e0.elapsed_time(e1) * 1e3
    total = sum(v[1] for v in agg.values())
    lines = [f"# per-layer device time, {a.model}@{a.width}x{a.height} {a.precision} batch {a.batch}, one eager forward ({len(m.ops.records)} op calls, "
             f"{total / 1e3:.2f} ms summed)", "", "| op:layer | first tensor shape | calls | total us | share | avg us |", "|---|---|---:|---:|---:|---:|"]
    for (label, shape), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        lines.append(f"| `{label}` | {shape} | {n} | {us:.0f} | {100 * us / total:.1f}% | {us / n:.1f} |")
    text = "\n".join(lines) + "\n"
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()
