#!/usr/bin/env python
"""Diagnostic (GPU): how far is fp16 from fp32 for (a) this repo's fp16 mode and (b) the reference algorithm under
torch.autocast(fp16) (the oracle run the way run_stereo_matching runs the reference, model_utils.py:76)?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import s2m2_oracle as O  # noqa: E402
from s2m2_b200.model import S2M2  # noqa: E402
from s2m2_b200.weights import random_state_dict, synthetic_pair  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda"
CASES = [("tiny neg", 32, 1, 2, 96, 64, False, 2, 3, 4), ("tiny pos", 32, 1, 1, 64, 96, True, 3, 0, 1),
         ("S 480x640", 128, 1, 1, 480, 640, True, 3, 0, 1), ("S neg 480x640", 128, 1, 1, 480, 640, False, 3, 0, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "L":
    CASES.append(("L 1024x1216", 256, 3, 1, 1024, 1216, True, 3, 0, 1))
for name, C, NTR, B, H, W, pos, it, ws, isd in CASES:
    sd = random_state_dict(C, NTR, ws)
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    l, r = synthetic_pair(B, H, W, isd)
    l, r = l.to(DEV), r.to(DEV)
    t32, t16 = {}, {}
    ref = O.forward(sdg, l, r, NTR, pos, it, t32)
    with torch.autocast("cuda", dtype=torch.float16):
        ac = O.forward(sdg, l, r, NTR, pos, it, t16)
    res = {}
    for prec in ("fp32", "fp16"):
        m = S2M2(C, 1, NTR, use_positivity=pos, refine_iter=it, precision=prec).eval()
        m.my_load_state_dict(sd)
        m = m.to(DEV)
        m.taps = {}
        out = m(l, r)
        res[prec] = (out, m.taps)
    def epe(a, b):
        return (a.float() - b.float()).abs().mean().item(), (a.float() - b.float()).abs().max().item()
    am32 = t32["disp_init.argmax"]
    print(f"== {name}: disp range [{ref[0].min().item():.2f},{ref[0].max().item():.2f}]")
    print("   reference-autocast-fp16 vs fp32 : EPE %.3e max %.3e | argmax flips %d / %d" %
          (*epe(ac[0], ref[0]), int((t16["disp_init.argmax"] != am32).sum()), am32.numel()))
    for prec in ("fp32", "fp16"):
        out, taps = res[prec]
        print("   s2m2_b200 %s vs oracle fp32     : EPE %.3e max %.3e | argmax flips %d | occ %.2e conf %.2e" %
              (prec, *epe(out[0], ref[0]), int((taps["disp_init.argmax"] != am32).sum()), epe(out[1], ref[1])[0], epe(out[2], ref[2])[0]))
