#!/usr/bin/env python
"""Host cost of one forward: the C++ schedule behind the handle-level C-ABI (s2m2b200_forward) against the Python schedule
(model.S2M2, eager ctypes calls) and CUDA-graph replay, same kernels.  L@1216x1024 by default.

    python tools/bench_host_forward.py [--model L --height 1024 --width 1216 --batches 1,8 --steps 10]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, n):
    """(ms per forward back to back on the device, ms per forward with a synchronize after each — the reference's protocol,
    model_utils.py:79-83 —, host ms to enqueue one forward into an idle stream)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t_host = t_sync = 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t_host += (t1 - t0) * 1e3 / n
        t_sync += (time.perf_counter() - t0) * 1e3 / n
    return ms, t_sync, t_host


KEYS = ("ms_per_forward", "ms_per_forward_sync_each", "host_enqueue_ms")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="L")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1216)
    ap.add_argument("--batches", default="1,8")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("-o", "--out", default=None)
    a = ap.parse_args()
    from s2m2_b200.host import HostForward
    from s2m2_b200.model import S2M2
    from s2m2_b200.weights import MODEL_TABLE, random_state_dict, synthetic_pair
    dev = torch.device("cuda", 0)
    C, NTR = MODEL_TABLE[a.model]
    sd = random_state_dict(C, NTR, seed=0)
    m = S2M2(C, 1, NTR, use_positivity=True, refine_iter=3, precision="fp16").eval()
    m.load_state_dict(sd)
    m = m.to(dev)
    h = HostForward(C, NTR, use_positivity=True, refine_iter=3, precision="fp16").load_state_dict(sd)
    hg = HostForward(C, NTR, use_positivity=True, refine_iter=3, precision="fp16", graph=True).load_state_dict(sd)
    rows = []
    for B in [int(b) for b in a.batches.split(",")]:
        left, right = [t.to(dev) for t in synthetic_pair(B, a.height, a.width, 1)]
        row = {"batch": B, "workspace_GB": h.workspace_bytes(B, a.height, a.width) / 1e9}
        with torch.inference_mode():
            m.auto_graph = False
            row["python_schedule_eager"] = dict(zip(KEYS, timed(lambda: m(left, right), a.steps)))
            row["cxx_schedule_eager"] = dict(zip(KEYS, timed(lambda: h(left, right), a.steps)))
            row["cxx_schedule_graph_mode"] = dict(zip(KEYS, timed(lambda: hg(left, right), a.steps)))
            m.auto_graph = True
            m._graphs = {}
            row["module_graph_replay"] = dict(zip(KEYS, timed(lambda: m(left, right), a.steps)))
            d0 = m(left, right)[0]
            d1 = h(left, right)[0]
            row["bit_identical"] = bool(torch.equal(d0, d1))
            m._graphs = {}
        torch.cuda.empty_cache()
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"model": a.model, "height": a.height, "width": a.width, "gpu": torch.cuda.get_device_name(0), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
