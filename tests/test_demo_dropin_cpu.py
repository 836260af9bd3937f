"""The reference's unmodified demo/visualize_2d_simple.py driven through tools/run_reference_demo.py.

Runs only where the reference tree exists (the build container).  The demo's own call chain — load_model
(model_utils.py:10-48), image_pad / run_stereo_matching / image_crop (model_utils.py:50-95) — is exercised
unchanged against the drop-in module.  No GPU here, so (test-only) the module gets the plain-torch op backend
and torch.cuda.Event / synchronize are shimmed; on a GPU box the launcher runs as is.
"""
import os
import sys
import time

import pytest
import torch

from conftest import ROOT

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "demo")), reason="reference tree not present")
def test_unmodified_demo_runs_with_dropin_module(tmp_path, monkeypatch, capsys):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_reference_demo as launcher
    from s2m2_b200 import model as model_mod
    from torch_ops import TorchOps
    calls = {"n": 0}

    class CpuBacked(model_mod.S2M2):        # test-only: the product class itself has no CPU path
        def __init__(self, *a, **k):
            super().__init__(*a, precision="fp32", **k)
            self.ops = TorchOps()

        def forward(self, l, r):
            calls["n"] += 1
            if calls["n"] > 1:               # the demo calls forward 6x at 1088x800; reuse the first result on CPU
                return self._last
            self._last = super().forward(l, r)
            return self._last

    monkeypatch.setattr(model_mod, "S2M2", CpuBacked)
    # the launcher rebinds s2m2.core.model.s2m2.S2M2 for the process: undo it after this test
    sys.path.insert(0, os.path.join(REF, "src"))
    import s2m2.core.model.s2m2 as ref_mod
    monkeypatch.setattr(ref_mod, "S2M2", ref_mod.S2M2)
    for k in [k for k in sys.modules if k.startswith("s2m2.core.utils")]:
        monkeypatch.delitem(sys.modules, k, raising=False)

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self):
            self.t = time.time()

        def elapsed_time(self, other):
            return max((other.t - self.t) * 1e3, 1e-3)
    if not torch.cuda.is_available():
        monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
        monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
        monkeypatch.setattr(torch.cuda, "manual_seed", lambda *a, **k: None)
    launcher.main(["--reference", REF, "--scratch", str(tmp_path), "--", "--model_type", "S", "--num_refine", "1"])
    out = capsys.readouterr().out
    assert "Model loaded" in out
    assert "cropped image size: img_height(1088), img_width(800)" in out
    assert "FPS:" in out
    assert calls["n"] == 6
    for k in [k for k in sys.modules if k.startswith("s2m2.core.utils")]:
        monkeypatch.delitem(sys.modules, k, raising=False)
