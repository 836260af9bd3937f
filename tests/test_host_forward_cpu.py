"""The C++ host forward (csrc/host_forward.cu, handle-level C-ABI) against the Python schedule (model.S2M2), on the CPU:

* the packed blob equals model._pack()'s tensors byte for byte (every layer, both precisions);
* with the op library in plan mode (nothing is launched; a box without a GPU plans for the sm_100a target) the two hosts make the
  same sequence of library calls with the same shapes, strides, weights, flags and buffer aliasing — and the workspace the C++
  host asks for covers its own allocations.
The outputs themselves are compared bit for bit on the GPU in tests/test_host_forward_gpu.py.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from s2m2_b200.host import HostForward
from s2m2_b200.model import ConvW, S2M2
from s2m2_b200.ops import load_library
from s2m2_b200.weights import random_state_dict

from plan_ops import PlanOps, canonical, check_arena, note_packed, trace_begin, trace_end

torch.set_num_threads(4)


def _build(Cc, NTR, prec, **kw):
    sd = random_state_dict(Cc, NTR, seed=3)
    m = S2M2(Cc, 1, NTR, precision=prec, **kw).eval()
    m.load_state_dict(sd)
    h = HostForward(Cc, NTR, precision=prec, **kw).load_state_dict(sd)
    return m, h


def _packed_tensors(m):
    out = {}

    def walk(o, path):
        if isinstance(o, ConvW):
            out["w:" + o.name] = o.w
            if o.b is not None:
                out["b:" + o.name] = o.b
        elif isinstance(o, dict):
            for k, v in o.items():
                walk(v, f"{path}.{k}" if path else str(k))
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                walk(v, f"{path}.{i}")
        elif isinstance(o, torch.Tensor):
            out[path] = o
    walk(m._packed, "")
    return out


@pytest.mark.parametrize("Cc,NTR,prec", [(64, 1, "fp16"), (64, 1, "fp32"), (96, 2, "fp16")])
def test_packed_blob_equals_python_packing(Cc, NTR, prec):
    m, h = _build(Cc, NTR, prec, use_positivity=True)
    m._ensure_packed(torch.device("cpu"))
    ref = _packed_tensors(m)
    ents = h.packed_entries()
    assert ents and not ents[0][4]                      # no device here: the blob stays on the host
    names = set()
    for name, ptr, nbytes, dt, _ in ents:
        names.add(name)
        t = ref[name]
        assert t.dtype == (torch.float16 if dt == 1 else torch.float32), name
        assert t.numel() * t.element_size() == nbytes, (name, tuple(t.shape), nbytes)
        assert C.string_at(ptr, nbytes) == t.contiguous().numpy().tobytes(), name
    # every packed tensor of the Python side is in the blob, except the tiny host-side tables that travel as kernel parameters
    host_side = {"bb.w0", "bb.b0", "bb.w1", "bb.b1", "u1.wd", "u1.bd", "u1.wr", "u1.br"}
    assert set(ref) - names == host_side


def test_missing_and_misshaped_parameters_are_reported():
    sd = random_state_dict(64, 1, seed=3)
    h = HostForward(64, 1)
    bad = dict(sd)
    del bad["refiner.gru.convz1.weight"]
    with pytest.raises(RuntimeError, match="refiner.gru.convz1.weight"):
        h.load_state_dict(bad)
    h = HostForward(64, 1)
    bad = dict(sd)
    bad["cnn_backbone.conv0.0.weight"] = torch.zeros(16, 4, 1, 1)
    with pytest.raises(RuntimeError, match="cnn_backbone.conv0.0.weight"):
        h.load_state_dict(bad)
    with pytest.raises(RuntimeError, match="multiples of 32"):
        HostForward(64, 1).load_state_dict(sd).workspace_bytes(1, 100, 128)


CASES = [
    # C, NTR, precision, B, H, W, dtype, crop, extra kwargs
    (64, 1, "fp16", 1, 128, 192, torch.float32, None, dict(use_positivity=True)),
    (64, 1, "fp32", 1, 96, 160, torch.float32, None, dict(use_positivity=False)),
    (64, 1, "fp16", 2, 128, 192, torch.uint8, (100, 180), dict(use_positivity=True, refine_iter=2)),
    (64, 1, "fp16", 1, 96, 160, torch.float32, None, dict(output_upsample=True)),
    (128, 1, "fp16", 2, 480, 640, torch.float32, None, dict(use_positivity=True)),          # S: fused token kernels above their thresholds
    (256, 3, "fp16", 1, 1024, 1216, torch.float32, None, dict(use_positivity=True)),        # L, the benchmarked configuration
    (192, 2, "fp16", 1, 256, 384, torch.float32, None, dict(use_positivity=True)),          # M: head dims 192 / 96 / 48 / 48
    (384, 3, "fp16", 1, 512, 640, torch.uint8, (500, 620), dict(use_positivity=True)),      # XL channels: the wide-C kernel variants
]


@pytest.mark.parametrize("Cc,NTR,prec,B,H,W,dt,crop,kw", CASES)
def test_host_forward_makes_the_same_library_calls_as_the_python_schedule(Cc, NTR, prec, B, H, W, dt, crop, kw):
    lib = load_library()
    m, h = _build(Cc, NTR, prec, **kw)
    # ---- Python schedule in plan mode
    ops = PlanOps()
    try:
        m.ops = ops
        m._ensure_packed(torch.device("cpu"))
        img0 = torch.empty((B, 3, H, W), dtype=dt)
        img1 = torch.empty((B, 3, H, W), dtype=dt)
        trace_begin(lib)
        note_packed(ops, m)
        ops.note(f"static in0 0x{img0.data_ptr():x} {img0.numel() * img0.element_size()}")
        ops.note(f"static in1 0x{img1.data_ptr():x} {img1.numel() * img1.element_size()}")
        m(img0, img1, crop_hw=crop)
        py = canonical(trace_end(lib))
    finally:
        ops.close()
        lib.s2m2b200_trace_end(None, C.c_int64(0))
    # ---- C++ schedule in plan mode (workspace_bytes walks it)
    trace_begin(lib)
    try:
        need = h.workspace_bytes(B, H, W, dt == torch.uint8, crop)
        raw = trace_end(lib)
    finally:
        lib.s2m2b200_trace_end(None, C.c_int64(0))
    cc = canonical(raw)
    assert len(py) > 100
    for i, (a, b) in enumerate(zip(py, cc)):
        assert a == b, f"line {i}:\n  python: {a}\n  c++   : {b}"
    assert len(py) == len(cc)
    assert "?ptr" not in "\n".join(cc)
    # the arena: every block inside the planned workspace, no two live blocks overlapping, everything released, and the plan is tight
    # (what it reports is the high-water mark, rounded to the 256-byte block granularity)
    peak = check_arena(raw, 0x100000000000, need)
    assert 0 < need - peak < 256 or need == peak
