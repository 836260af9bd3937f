"""The C-ABI shared library loads and exports every symbol include/s2m2b200.h declares (no GPU needed)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "s2m2b200.h")).read()
    return sorted(set(re.findall(r"\b(s2m2b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from s2m2_b200 import ops
    if not os.path.exists(ops.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ops.load_library()
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/s2m2b200.h but not exported"
    assert sorted(ops.EXPORTS) == names
    assert b"sm_100a" in lib.s2m2b200_version()


def test_pe_table_host_function_matches_oracle():
    import ctypes as C
    import torch
    from oracle.s2m2_oracle import pe_tables
    from s2m2_b200 import ops
    lib = ops.load_library()
    for n in (2, 15, 38):
        host = torch.empty((2 * n + 1, 16))
        assert lib.s2m2b200_pe_table(n, C.c_void_p(host.data_ptr())) == 0
        ref, _ = pe_tables(3, n)
        assert (host - ref).abs().max() < 2e-6


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from s2m2_b200.model import S2M2
    m = S2M2(32, 1, 1, True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64))


def test_module_surface_matches_nn_module_conventions():
    """state_dict(prefix=..., destination=...) and load_state_dict's _IncompatibleKeys, as wrappers (EMA, DDP-style utilities,
    submodule-aware loaders) expect them."""
    from collections import OrderedDict

    import torch

    from s2m2_b200.model import S2M2
    from s2m2_b200.weights import random_state_dict
    m = S2M2(32, 1, 1)
    sd = random_state_dict(32, 1, 0)
    res = m.load_state_dict(sd)
    assert res.missing_keys == [] and res.unexpected_keys == []
    part = dict(sd)
    dropped = part.pop("ctx_feat.0.weight")
    part["not.a.key"] = torch.zeros(1)
    res = m.load_state_dict(part, strict=False)
    assert res.missing_keys == ["ctx_feat.0.weight"] and res.unexpected_keys == ["not.a.key"]
    out = m.state_dict(prefix="model.", destination=OrderedDict(extra=1))
    assert out["extra"] == 1 and torch.equal(out["model.ctx_feat.0.weight"], dropped)
    assert not out["model.ctx_feat.0.weight"].requires_grad
    assert isinstance(m.state_dict(keep_vars=True)["ctx_feat.0.weight"], torch.nn.Parameter)
    assert list(m.state_dict().keys()) == list(sd.keys())


def test_runner_rejects_silently_lossy_default():
    """PipelinedRunner.run() must not accept more batches than its two rotating default output sets."""
    import inspect

    from s2m2_b200 import runtime
    src = inspect.getsource(runtime.PipelinedRunner.run)
    assert "host_out is None and n > 2" in src
