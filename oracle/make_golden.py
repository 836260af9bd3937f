"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

For each case: synthetic weights (s2m2_b200.weights.random_state_dict — deterministic, so they are
regenerated rather than stored) are loaded into the unmodified reference ``S2M2`` module, the
synthetic pair is pushed through ``model(left, right)`` in fp32 on CPU, and the three outputs plus
stage taps (forward hooks on reference sub-modules) are stored.  The reference ships no golden
vectors of its own, so these files are the pin for ``oracle/s2m2_oracle.py``.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from s2m2.core.model.s2m2 import S2M2  # noqa: E402  (the real reference)
from s2m2_b200.weights import random_state_dict, synthetic_pair  # noqa: E402

CASES = {
    # name: (C, NTR, B, H, W, use_positivity, refine_iter, weight_seed, input_seed, gain)
    "tiny_c32_64x96": (32, 1, 1, 64, 96, True, 3, 0, 1, 1.0),
    "tiny_c32_b2_neg_96x64": (32, 1, 2, 96, 64, False, 2, 3, 4, 1.0),
    "tiny_c64_ntr2_64x128": (64, 2, 1, 64, 128, True, 1, 5, 6, 1.0),
    "s_c128_96x160": (128, 1, 1, 96, 160, True, 3, 0, 1, 1.0),
    "tiny_c32_up2_64x96": (32, 1, 1, 64, 96, True, 2, 7, 8, 1.0, True),     # output_upsample=True (s2m2.py:123-127,194-195)
}


def run_case(name, cfg):
    C, NTR, B, H, W, pos, it, wseed, iseed, gain = cfg[:10]
    up2 = bool(cfg[10]) if len(cfg) > 10 else False
    sd = random_state_dict(C, NTR, wseed, gain)
    m = S2M2(C, 1, NTR, pos, up2, it).eval()
    m.load_state_dict(sd, strict=True)
    left, right = synthetic_pair(B, H, W, iseed)
    taps = {}

    def hook(key):
        def fn(mod, inp, out):
            outs = out if isinstance(out, (tuple, list)) else (out,)
            n = sum(1 for k in taps if k.startswith(key + "#"))  # call counter (refiner fires refine_iter times)
            for i, o in enumerate(outs):
                taps[f"{key}#{n // len(outs)}.{i}"] = o.detach().clone().numpy()
        return fn

    for key in ("cnn_backbone", "feat_pyramid", "transformer", "disp_init", "global_refiner",
                "feat_fusion_layer", "ctx_feat", "refiner", "upsample_mask_4x_refine", "upsample_mask_1x"):
        getattr(m, key).register_forward_hook(hook(key))
    with torch.inference_mode():
        d, o, c = m(left, right)
    out = {"disp": d.numpy(), "occ": o.numpy(), "conf": c.numpy(),
           "cfg": np.array([C, NTR, B, H, W, int(pos), it, wseed, iseed], dtype=np.int64),
           "gain": np.array([gain], dtype=np.float64), "up2": np.array([int(up2)], dtype=np.int64)}
    keep = ("cnn_backbone#0.0", "feat_pyramid#0.0", "feat_pyramid#0.3", "transformer#0.0",
            "disp_init#0.0", "disp_init#0.1", "disp_init#0.2", "disp_init#0.3", "global_refiner#0.0",
            "ctx_feat#0.0", "upsample_mask_4x_refine#0.0")
    for k in keep:
        if C <= 64 or taps[k].size <= 64 * 1024:   # keep fixtures small: big feature taps only for tiny models
            out["tap:" + k] = taps[k]
    for i in range(it):
        for j, nm in enumerate(("hidden", "disp", "conf", "occ")):
            if nm != "hidden":
                out[f"tap:refiner#{i}.{nm}"] = taps[f"refiner#{i}.{j}"]
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB",
          "disp range", float(d.min()), float(d.max()))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if not only or name in only:
            run_case(name, cfg)
