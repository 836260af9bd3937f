"""Test infrastructure: walk ``model.S2M2``'s schedule on the CPU with the op library in plan mode and record the op trace.

``PlanOps`` is ``ops.CudaOps`` on CPU tensors: every method makes the same call into libs2m2b200.so, which — with
``s2m2b200_set_plan_mode(1)`` — runs its host-side checks and dispatch (which kernel takes this layer, which return code the
caller gets) and launches nothing.  With ``s2m2b200_trace_begin`` open each call leaves one line; allocations are noted as
``alloc <ptr> <bytes>`` and long-lived buffers as ``static <name> <ptr> <bytes>``.  ``canonical`` rewrites the addresses in a trace
as ``a<serial>+<offset>`` / ``<name>+<offset>`` so that the traces of two hosts (this Python schedule, the C++ one in
csrc/host_forward.cu) can be compared line by line.
"""
import bisect
import ctypes as C
import re

import torch

from s2m2_b200.model import ConvW
from s2m2_b200.ops import CudaOps


class PlanOps(CudaOps):
    def _check_device(self):
        pass

    def __init__(self):
        super().__init__()
        self.lib.s2m2b200_set_plan_mode(1)

    def close(self):
        self.lib.s2m2b200_set_plan_mode(0)

    @staticmethod
    def _stream():
        return None

    def note(self, text):
        self.lib.s2m2b200_trace_note(text.encode())

    def _aligned(self, shape, dtype):
        """256-byte aligned like the CUDA allocators (the ops' alignment checks decide which kernel takes a layer)."""
        numel = 1
        for d in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
            numel *= int(d)
        nbytes = max(numel * torch.empty((), dtype=dtype).element_size(), 1)
        base = torch.empty(nbytes + 256, dtype=torch.uint8)
        off = (-base.data_ptr()) % 256
        t = base[off:off + nbytes].view(dtype).view(shape)
        self.note(f"alloc 0x{t.data_ptr():x} {nbytes}")
        return t

    def empty(self, shape, dtype, device):
        return self._aligned(tuple(shape), dtype)

    def zeros(self, shape, dtype, device):
        return self._aligned(tuple(shape), dtype)          # contents are never read in plan mode

    def _scratch(self, numel, dtype, device, zero=False):
        if dtype == torch.float16 and numel <= 8:          # ops.attention's cached token buffer: a named region (the C++ host
            t = torch.empty(64, dtype=dtype)               # keeps it with the shape's constants)
            self.note(f"static attn_work 0x{t.data_ptr():x} {t.numel() * 2}")
            return t
        return self._aligned((numel,), dtype)

    def pe_tables(self, h, w, device):
        fresh = (h, w, str(device)) not in self._pe
        tabs = super().pe_tables(h, w, device)
        if fresh:
            self.note(f"static pe_x 0x{tabs[0].data_ptr():x} {tabs[0].numel() * 4}")
            self.note(f"static pe_y 0x{tabs[1].data_ptr():x} {tabs[1].numel() * 4}")
            if tabs[4] is not None:
                self.note(f"static pe_ind 0x{tabs[4].data_ptr():x} {tabs[4].numel() * 2}")
        return tabs


def note_packed(ops, model):
    """Name every packed weight region of ``model`` (after ``_ensure_packed``) in the open trace."""
    seen = set()

    def walk(o, path):
        if isinstance(o, ConvW):
            if id(o) in seen:
                return
            seen.add(id(o))
            ops.note(f"static w:{o.name} 0x{o.w.data_ptr():x} {o.w.numel() * o.w.element_size()}")
            if o.b is not None:
                ops.note(f"static b:{o.name} 0x{o.b.data_ptr():x} {o.b.numel() * 4}")
        elif isinstance(o, dict):
            for k, v in o.items():
                walk(v, f"{path}.{k}" if path else str(k))
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                walk(v, f"{path}.{i}")
        elif isinstance(o, torch.Tensor):
            ops.note(f"static {path} 0x{o.data_ptr():x} {o.numel() * o.element_size()}")

    P = model._packed
    walk({k: v for k, v in P.items() if k not in ("bb", "u1")}, "")
    bb, u1 = P["bb"], P["u1"]
    for k in ("d0", "d2", "c20", "c22", "down"):
        walk(bb[k], "")
    for k in ("gn_w", "gn_b", "w0", "b0", "w1", "b1"):
        walk(bb[k], f"bb.{k}")
    for k in ("ctx", "c0", "c2"):
        walk(u1[k], "")
    for k in ("wd", "bd", "wr", "br"):
        walk(u1[k], f"u1.{k}")


def trace_begin(lib):
    lib.s2m2b200_trace_begin()


def trace_end(lib) -> str:
    lib.s2m2b200_trace_end.restype = C.c_int64
    n = int(lib.s2m2b200_trace_end(None, C.c_int64(0)))    # stops the trace and returns its length; the text stays until the next _begin
    buf = C.create_string_buffer(n + 1)
    lib.s2m2b200_trace_end(buf, C.c_int64(n + 1))
    return buf.value.decode()


_PTR = re.compile(r"0x[0-9a-f]+")


def canonical(trace: str):
    """Rewrite every address as ``<region>+<offset>``: ``a<k>`` for the k-th allocation, the given name for a static region."""
    starts, regions = [], {}       # sorted starts; start -> (end, label)
    serial = 0
    out = []

    def add(start, nbytes, label):
        end = start + max(nbytes, 1)
        # drop stale regions this one overlaps (the allocator reused their memory)
        i = bisect.bisect_left(starts, start)
        while i > 0 and regions[starts[i - 1]][0] > start:
            i -= 1
        while i < len(starts) and starts[i] < end:
            if regions[starts[i]][0] > start:
                del regions[starts[i]]
                del starts[i]
            else:
                i += 1
        bisect.insort(starts, start)
        regions[start] = (end, label)

    def name_of(m):
        p = int(m.group(0), 16)
        i = bisect.bisect_right(starts, p) - 1
        if i >= 0:
            end, label = regions[starts[i]]
            if p < end:
                return f"{label}+{p - starts[i]}"
        return "?ptr"

    for line in trace.splitlines():
        parts = line.split()
        if parts and parts[0] == "alloc":
            add(int(parts[1], 16), int(parts[2]), f"a{serial}")
            out.append(f"alloc a{serial} {parts[2]}")
            serial += 1
        elif parts and parts[0] == "static":
            add(int(parts[2], 16), int(parts[3]), parts[1])
        elif parts and parts[0] == "free":
            pass                                   # lifetimes: only the C++ arena reports them (check_arena below)
        else:
            out.append(_PTR.sub(name_of, line))
    return out


def check_arena(trace: str, base: int, need: int):
    """The C++ host's workspace arena, from its raw trace: every block inside [base, base + need), and no two LIVE blocks overlap
    (``alloc`` / ``free`` lines; blocks outside the arena's address range — the caller's output buffer — are ignored)."""
    live = {}
    peak = 0
    for line in trace.splitlines():
        p = line.split()
        if not p:
            continue
        if p[0] == "alloc":
            a, n = int(p[1], 16), max(int(p[2]), 1)
            if not (base <= a < base + (1 << 44)):
                continue
            assert a + n <= base + need, f"block {a - base}+{n} beyond the planned workspace {need}"
            for b, m in live.items():
                assert a + n <= b or b + m <= a, f"live blocks overlap: {a - base}+{n} and {b - base}+{m}"
            live[a] = n
            peak = max(peak, a + n - base)
        elif p[0] == "free":
            assert int(p[1], 16) in live, "free of an unknown block"
            del live[int(p[1], 16)]
    assert not live, f"{len(live)} blocks never released"
    return peak
