import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    C, NTR, B, H, W, pos, it, wseed, iseed = [int(v) for v in z["cfg"]]
    cfg = dict(C=C, NTR=NTR, B=B, H=H, W=W, pos=bool(pos), it=it, wseed=wseed, iseed=iseed, gain=float(z["gain"][0]),
               up2=bool(z["up2"][0]) if "up2" in z.files else False)
    data = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("cfg", "gain", "up2")}
    return cfg, data


GOLDEN_CASES = ["tiny_c32_64x96", "tiny_c32_b2_neg_96x64", "tiny_c64_ntr2_64x128", "s_c128_96x160"]
GOLDEN_UP2 = "tiny_c32_up2_64x96"
