#!/usr/bin/env python
"""One steady-state forward between cudaProfilerStart / Stop — the target of the ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/launch_list.py [--batch 8] [--model L --height 1024 --width 1216]
    python tools/summarize_launches.py gpurun_out/launches.csv -o profiles/rN_launches.md
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2m2_b200.model import S2M2  # noqa: E402
from s2m2_b200.weights import MODEL_TABLE, random_state_dict, synthetic_pair  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="L")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1216)
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    C, NTR = MODEL_TABLE[a.model]
    m = S2M2(C, 1, NTR, use_positivity=True, refine_iter=3, precision="fp16").eval()
    m.my_load_state_dict(random_state_dict(C, NTR, 0))
    m = m.cuda()
    m.auto_graph = False
    l, r = synthetic_pair(a.batch, a.height, a.width, 1)
    l, r = l.cuda(), r.cuda()
    with torch.inference_mode():
        for _ in range(2):
            m(l, r)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        m(l, r)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("launch list forward done; launches per forward:", m.ops.launch_count() // 3)


if __name__ == "__main__":
    main()
