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
