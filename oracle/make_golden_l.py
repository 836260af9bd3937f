"""Generate tests/golden/l_c256_1216x1024_raw.npz: the REAL reference at the BASELINE headline config (L = CH256/NTR3,
1216x1024, refine_iter=3, un-softened BASELINE.md §4 inputs), compacted so that the fixture stays a few MB.

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference; ~1 min on 8 cores):

    python oracle/make_golden_l.py

The unmodified reference ``S2M2`` (imported from /root/reference/src) is loaded with the deterministic synthetic weights
``random_state_dict(256, 3, 0)`` and run in fp32 on CPU.  Stored:
  * the three full-resolution outputs sub-sampled by 4 in both directions (`disp_s4`, `occ_s4`, `conf_s4`) and their
    full-resolution means (`mean`) — a checksum of everything in between;
  * the integer contract (north_star "bit-exact for the argmax index and occlusion masks"):
    `argmax` = prob_max_ind (submodules.py:226, recomputed from the hooked cost volume with the reference's OWN
    ``_optimal_transport``), `conf_mask` = conf > .2 (refinenet.py:65), `occ_mask{i}` = x - disp >= 0 after refine iteration i
    (s2m2.py:177-180, from the hooked refiner output with the same expression), bit-packed;
  * the 1/4-resolution maps of DispInit and of every refine iteration (`di_*`, `it{i}_disp`).
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

NAME = "l_c256_1216x1024_raw"
C, NTR, B, H, W, IT, WSEED, ISEED = 256, 3, 1, 1024, 1216, 3, 0, 1


def main():
    torch.set_num_threads(os.cpu_count())
    m = S2M2(C, 1, NTR, True, False, IT).eval()
    m.load_state_dict(random_state_dict(C, NTR, WSEED), strict=True)
    left, right = synthetic_pair(B, H, W, ISEED, smooth=False)
    taps = {}

    def hook(key):
        def fn(mod, inp, out):
            n = sum(1 for k in taps if k.startswith(key + "#"))
            taps[f"{key}#{n}"] = [o.detach().clone() for o in (out if isinstance(out, (tuple, list)) else (out,))]
        return fn
    m.disp_init.register_forward_hook(hook("disp_init"))
    m.refiner.register_forward_hook(hook("refiner"))
    with torch.inference_mode():
        d, o, c = m(left, right)
        di_disp, di_conf, di_occ, cv = taps["disp_init#0"]
        w4 = cv.shape[-1]
        mask = torch.triu(torch.ones((w4, w4), dtype=torch.bool), diagonal=1)           # submodules.py:211-212
        prob = m.disp_init._optimal_transport(cv.masked_fill(mask, -1e4)).masked_fill(mask, 0)   # :221-223, the reference's own code
        argmax = prob.argmax(dim=3)                                                        # :226
        xs = torch.arange(w4, dtype=torch.float32).reshape(1, 1, 1, -1)
        out = {"cfg": np.array([C, NTR, B, H, W, 1, IT, WSEED, ISEED], dtype=np.int64), "smooth": np.array([0]), "stride": np.array([4]),
               "disp_s4": d[..., ::4, ::4].numpy(), "occ_s4": o[..., ::4, ::4].numpy(), "conf_s4": c[..., ::4, ::4].numpy(),
               "mean": np.array([d.double().mean(), o.double().mean(), c.double().mean()]),
               "argmax": argmax.numpy().astype(np.int16), "conf_mask": np.packbits((di_conf > 0.2).numpy()),
               "di_disp": di_disp.numpy(), "di_conf": di_conf.numpy(), "di_occ": di_occ.numpy()}
        for i in range(IT):
            disp_i = taps[f"refiner#{i}"][1].clamp(min=0)                                 # s2m2.py:177-178
            out[f"occ_mask{i}"] = np.packbits(torch.ge(xs - disp_i, 0).numpy())            # s2m2.py:179
            out[f"it{i}_disp"] = disp_i.numpy()
    path = os.path.join(ROOT, "tests", "golden", NAME + ".npz")
    np.savez_compressed(path, **out)
    print(NAME, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB", "disp range", float(d.min()), float(d.max()),
          "argmax margin check: prob top1 mean", float(prob.max(dim=3).values.mean()))


if __name__ == "__main__":
    main()
