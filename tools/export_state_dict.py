#!/usr/bin/env python
"""Write a state dict (the reference's CH{C}NTR{N}.pth, or synthetic weights) as the flat binary file examples/host_forward.c reads:
"S2M2SD01", int32 count, then per entry int32 name_len, name, int32 ndim, int64 shape[ndim], float32 data.

    python tools/export_state_dict.py --model S -o /tmp/s.bin              # synthetic (random_state_dict, seed 0)
    python tools/export_state_dict.py --checkpoint CH256NTR3.pth -o l.bin  # a reference checkpoint (model_utils.py:30-43)
"""
import argparse
import os
import struct
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_state_dict(sd, path):
    with open(path, "wb") as f:
        f.write(b"S2M2SD01")
        f.write(struct.pack("<i", len(sd)))
        for name, t in sd.items():
            a = t.detach().cpu().float().contiguous().numpy()
            nb = name.encode()
            f.write(struct.pack("<i", len(nb)))
            f.write(nb)
            f.write(struct.pack("<i", a.ndim))
            f.write(struct.pack(f"<{a.ndim}q", *a.shape))
            f.write(a.astype("<f4").tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="S")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("-o", "--out", required=True)
    a = ap.parse_args()
    if a.checkpoint:
        sd = torch.load(a.checkpoint, map_location="cpu", weights_only=True)
        sd = sd.get("state_dict", sd)
    else:
        from s2m2_b200.weights import MODEL_TABLE, random_state_dict
        sd = random_state_dict(*MODEL_TABLE[a.model], seed=a.seed)
    write_state_dict(sd, a.out)
    print(f"{a.out}: {len(sd)} entries")


if __name__ == "__main__":
    main()
