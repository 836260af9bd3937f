"""ctypes binding of the s2m2b200 C-ABI (include/s2m2b200.h) — the op backend used by ``model.S2M2``.

Thin by design: torch is used only for device memory (caching allocator) and the current stream; every
op is one call into ``libs2m2b200.so`` with raw pointers.  There is deliberately no fallback: if the
library is missing or there is no CUDA device, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2m2b200.so")

F32, F16 = 0, 1
_DT = {torch.float32: F32, torch.float16: F16}


class S2m2Tensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("C", C.c_int32), ("ld", C.c_int64)]


class S2m2Conv(C.Structure):
    _fields_ = [("x", S2m2Tensor), ("x2", S2m2Tensor), ("y", S2m2Tensor), ("res", S2m2Tensor),
                ("w", C.c_void_p), ("w_dtype", C.c_int32), ("bias", C.c_void_p),
                ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32),
                ("Cout", C.c_int32), ("act", C.c_int32), ("shuffle2x2", C.c_int32), ("batch", C.c_int32),
                ("x_bs", C.c_int64), ("w_bs", C.c_int64), ("y_bs", C.c_int64), ("impl", C.c_int32),
                ("res2", S2m2Tensor), ("emode", C.c_int32), ("res_from", C.c_int32)]


_lib = None


def load_library():
    """dlopen libs2m2b200.so (built in-tree by ``__graft_entry__.build()`` / ``make -C s2m2_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU / PyTorch fallback for the S2M2 forward)")
    lib = C.CDLL(LIB_PATH)
    lib.s2m2b200_last_error.restype = C.c_char_p
    lib.s2m2b200_version.restype = C.c_char_p
    lib.s2m2b200_launch_count.restype = C.c_int64
    lib.s2m2b200_attention_workspace_bytes.restype = C.c_int64
    lib.s2m2b200_pe_indicator_bytes.restype = C.c_int64
    lib.s2m2b200_row_corr_workspace_bytes.restype = C.c_int64
    lib.s2m2b200_trace_end.restype = C.c_int64
    lib.s2m2b200_trace_note.restype = None
    _lib = lib
    return lib


EXPORTS = [
    "s2m2b200_conv2d", "s2m2b200_ffn", "s2m2b200_mlp2", "s2m2b200_ln_linear", "s2m2b200_attention", "s2m2b200_attention_workspace_bytes", "s2m2b200_pe_table", "s2m2b200_pe_indicator", "s2m2b200_pe_indicator_bytes", "s2m2b200_row_corr", "s2m2b200_row_corr_workspace_bytes", "s2m2b200_layernorm", "s2m2b200_groupnorm",
    "s2m2b200_avgpool2", "s2m2b200_upsample2x", "s2m2b200_gate_blend", "s2m2b200_copy_channels", "s2m2b200_unary",
    "s2m2b200_mul", "s2m2b200_gru_update", "s2m2b200_preprocess", "s2m2b200_stem", "s2m2b200_mask1x_inputs", "s2m2b200_sinkhorn_softargmax", "s2m2b200_cv_lookup",
    "s2m2b200_global_refiner_prep", "s2m2b200_global_refiner_merge", "s2m2b200_local_refiner_prep",
    "s2m2b200_local_refiner_update", "s2m2b200_upsample4x_convex", "s2m2b200_filter1x", "s2m2b200_filter1x_crop", "s2m2b200_image_pad", "s2m2b200_remap_u8", "s2m2b200_last_error",
    "s2m2b200_preprocess_u8", "s2m2b200_stem_u8", "s2m2b200_confidence_score", "s2m2b200_disparity_to_points",
    "s2m2b200_version", "s2m2b200_launch_count", "s2m2b200_set_plan_mode",
    "s2m2b200_trace_begin", "s2m2b200_trace_note", "s2m2b200_trace_end",
    # the whole forward behind one handle (host side in C++; bound in s2m2_b200/host.py)
    "s2m2b200_create", "s2m2b200_set_param", "s2m2b200_finalize_params", "s2m2b200_workspace_bytes", "s2m2b200_forward", "s2m2b200_destroy", "s2m2b200_set_graph_mode", "s2m2b200_stereo_match", "s2m2b200_stereo_match_workspace_bytes",
    "s2m2b200_packed_count", "s2m2b200_packed_entry",
]


def tview(t: torch.Tensor) -> S2m2Tensor:
    """Describe an NHWC torch view (last-dim stride 1, pixels equally spaced) to the C side."""
    if t.dim() != 4 or t.dtype not in _DT:
        raise ValueError(f"expected a 4-D fp32/fp16 NHWC tensor, got {tuple(t.shape)} {t.dtype}")
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    ld = sw if w > 1 else (sh if h > 1 else (sn if n > 1 else c))
    ok = (sc == 1 or c == 1) and (w == 1 or sw == ld) and (h == 1 or sh == w * ld) and (n == 1 or sn == h * w * ld)
    if not ok:
        raise ValueError(f"tensor view is not a strided-NHWC view: shape {tuple(t.shape)} strides {t.stride()}")
    return S2m2Tensor(t.data_ptr(), _DT[t.dtype], n, h, w, c, ld)


_NULL_T = S2m2Tensor(None, 0, 0, 0, 0, 0, 0)


class CudaOps:
    """The product's op backend: every method enqueues hand-written CUDA kernels on the current stream."""

    def __init__(self, conv_impl: int = None):
        self._check_device()
        self.lib = load_library()
        if conv_impl is None:
            conv_impl = int(os.environ.get("S2M2_CONV_IMPL", "0"))
        self.conv_impl = conv_impl      # 0 auto, 1 force CUDA-core kernel, 2 force tcgen05 kernel
        self.ffn_fused = os.environ.get("S2M2_FFN_FUSED", "1") != "0"     # fused LN+FFN tcgen05 kernel
        self.ffn_min_tokens = int(os.environ.get("S2M2_FFN_MIN_TOKENS", "16384"))
        # C > 256 runs the one-tile-at-a-time kernel (the tile pair does not fit TMEM): it needs >= 2 tiles per SM to win
        self.ffn_min_tokens_wide = int(os.environ.get("S2M2_FFN_MIN_TOKENS_WIDE", "37888"))
        self.row_corr_fused = os.environ.get("S2M2_ROWCORR_TC", "1") != "0"   # tcgen05 split-fp16 cost volume with fused LayerNorm
        self.fuse_epilogues = os.environ.get("S2M2_FUSE_EPILOGUES", "1") != "0"   # gate blend / GRU steps as conv epilogue modes
        self.mlp2_fused = os.environ.get("S2M2_MLP2_FUSED", "1") != "0"   # fused 1x1 -> ReLU -> 1x1 (+res) of ConvBlock2D
        self.mlp2_min_tokens = int(os.environ.get("S2M2_MLP2_MIN_TOKENS", "16384"))
        self.mlp2_max_tokens = int(os.environ.get("S2M2_MLP2_MAX_TOKENS", "120000"))
        self.ln_linear_fused = os.environ.get("S2M2_LNL_FUSED", "1") != "0"  # fused LN + QKV projection tcgen05 kernel
        self.ln_linear_min_tokens = int(os.environ.get("S2M2_LNL_MIN_TOKENS", "16384"))
        self.attn_tc = os.environ.get("S2M2_ATTN_TC", "1") != "0"     # tcgen05 whole-row attention for the scan-line blocks
        self._pe = {}
        self._gn_work = {}
        self._attn_work = {}

    # -- helpers
    def _check_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("s2m2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")

    def _scratch(self, numel, dtype, device, zero=False):
        """Op-internal scratch / cached work buffers (one place, so that the plan-mode harness in tests/ sees them)."""
        return (torch.zeros if zero else torch.empty)(numel, dtype=dtype, device=device)

    def _chk(self, rc, name):
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {self.lib.s2m2b200_last_error().decode()}")

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launch_count(self) -> int:
        return int(self.lib.s2m2b200_launch_count())

    def empty(self, shape, dtype, device):
        return torch.empty(shape, dtype=dtype, device=device)

    def zeros(self, shape, dtype, device):
        return torch.zeros(shape, dtype=dtype, device=device)

    def pe_tables(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._pe:
            tabs = []
            for n in (w, h):
                host = torch.empty((2 * n + 1, 16), dtype=torch.float32)
                self._chk(self.lib.s2m2b200_pe_table(n, C.c_void_p(host.data_ptr())), "pe_table")
                tabs.append(host.to(device))
            ind = None
            if h + w <= 80:        # indicator matrix of the token grid: the tcgen05 PE attention takes the marginals of P as P . E
                nb = int(self.lib.s2m2b200_pe_indicator_bytes(h, w))
                hostb = torch.empty(nb // 2, dtype=torch.float16)
                self._chk(self.lib.s2m2b200_pe_indicator(h, w, C.c_void_p(hostb.data_ptr())), "pe_indicator")
                ind = hostb.to(device)
            self._pe[key] = (tabs[0], tabs[1], h, w, ind)
        return self._pe[key]

    # -- GEMM-class
    # epilogue modes of the tcgen05 conv (include/s2m2b200.h): the gate / GRU element-wise steps folded into the producing layer
    EMODE_ADD, EMODE_MUL, EMODE_BLEND, EMODE_GRU = 0, 1, 2, 3

    def conv_emodes(self, x) -> bool:
        """Are the fused gate / GRU epilogues available for activations like ``x``?  (fp16 tcgen05 path only.)"""
        return self.fuse_epilogues and x.dtype == torch.float16 and self.conv_impl != 1

    def conv2d(self, x, cw, act, out, res=None, x2=None, batch=1, x_bs=0, w_bs=0, y_bs=0, res2=None, emode=0, res_from=0):
        a = S2m2Conv()
        a.res2 = tview(res2) if res2 is not None else _NULL_T
        a.emode, a.res_from = emode, res_from
        a.x = tview(x)
        a.x2 = tview(x2) if x2 is not None else _NULL_T
        a.y = tview(out)
        a.res = tview(res) if res is not None else _NULL_T
        cin = x.shape[3] + (x2.shape[3] if x2 is not None else 0)
        if cin != cw.cin:
            raise ValueError(f"conv {cw.name}: input has {cin} channels, weights expect {cw.cin}")
        a.w = cw.w.data_ptr()
        a.w_dtype = _DT[cw.w.dtype]
        a.bias = cw.b.data_ptr() if cw.b is not None else None
        a.KH, a.KW, a.stride, a.pad_h, a.pad_w = cw.kh, cw.kw, cw.stride, cw.ph, cw.pw
        a.Cout, a.act, a.shuffle2x2 = cw.cout, act, 1 if cw.shuffle else 0
        a.batch, a.x_bs, a.w_bs, a.y_bs = batch, x_bs, w_bs, y_bs
        a.impl = self.conv_impl
        rc = self.lib.s2m2b200_conv2d(C.byref(a), self._stream())
        if rc == 1 and emode != 0:
            return False                 # fused epilogue mode not available for this layer: the caller runs the separate ops
        self._chk(rc, f"conv2d[{cw.name}]")
        return True

    def ffn(self, z, l0, l2):
        """Fused LN -> Linear -> GELU -> Linear -> +z in place; False if the fused kernel does not take this shape/dtype."""
        if not self.ffn_fused or l0.b is None or l2.b is None or l0.cin != l0.cout or l2.cin != l2.cout:
            return False
        # the fused kernel runs one CTA per 128-token tile: below ~2 tiles per SM (1/16 and 1/32 resolution at L) the three
        # separate launches, which also split the output channels over CTAs, are faster (measured, tools/bench_ffn.py)
        if z.shape[0] * z.shape[1] * z.shape[2] < (self.ffn_min_tokens if z.shape[3] <= 256 else self.ffn_min_tokens_wide):
            return False
        zt = tview(z)
        rc = self.lib.s2m2b200_ffn(C.byref(zt), C.byref(zt), C.c_void_p(l0.w.data_ptr()), C.c_void_p(l0.b.data_ptr()),
                                   C.c_void_p(l2.w.data_ptr()), C.c_void_p(l2.b.data_ptr()), _DT[l0.w.dtype], self._stream())
        if rc == 1:
            return False
        self._chk(rc, "ffn")
        return True

    def mlp2(self, x, res, out, l0, l2, act):
        """out = res + W2 . act(W0 . x + b0) + b2 in one tcgen05 kernel (ConvBlock2D's 1x1 branch, attentions.py:273-281);
        False if the fused kernel does not take this shape / dtype (the caller then runs the two conv2d launches)."""
        if not self.mlp2_fused or l0.b is None or l2.b is None or l0.cin != l0.cout or l2.cin != l2.cout or l0.kh != 1 or l2.kh != 1:
            return False
        ntok = x.shape[0] * x.shape[1] * x.shape[2]
        if ntok < self.mlp2_min_tokens or ntok > self.mlp2_max_tokens:     # large maps: two HBM-bound GEMMs are as fast (measured)
            return False
        rc = self.lib.s2m2b200_mlp2(C.byref(tview(x)), C.byref(tview(res)) if res is not None else None, C.byref(tview(out)),
                                    C.c_void_p(l0.w.data_ptr()), C.c_void_p(l0.b.data_ptr()), C.c_void_p(l2.w.data_ptr()),
                                    C.c_void_p(l2.b.data_ptr()), _DT[l0.w.dtype], 0, act, self._stream())
        if rc == 1:
            return False
        self._chk(rc, "mlp2")
        return True

    def ln_linear(self, z, lin, out):
        """out = Linear(LayerNorm(z)) in one tcgen05 kernel (LN(z) never leaves shared memory); False if not applicable."""
        if not self.ln_linear_fused or lin.kh != 1 or lin.kw != 1 or lin.shuffle:
            return False
        if z.shape[0] * z.shape[1] * z.shape[2] < self.ln_linear_min_tokens:
            return False
        rc = self.lib.s2m2b200_ln_linear(C.byref(tview(z)), C.byref(tview(out)), C.c_void_p(lin.w.data_ptr()),
                                         C.c_void_p(lin.b.data_ptr() if lin.b is not None else None), _DT[lin.w.dtype], self._stream())
        if rc == 1:
            return False
        self._chk(rc, "ln_linear")
        return True

    def row_corr(self, f, cv):
        """cv[b,y,i,j] = <f[b,y,i,:], f[B+b,y,j,:]>: one GEMM per scan-line (batched), fp32 (submodules.py:216-217)."""
        B, h, w, _ = cv.shape
        Cc = f.shape[3]
        if f.dtype != torch.float32 or not f.is_contiguous() or not cv.is_contiguous():
            raise ValueError("row_corr expects contiguous fp32 features / cost volume")
        a = S2m2Conv()
        a.x = S2m2Tensor(f.data_ptr(), F32, 1, 1, w, Cc, Cc)
        a.x2 = _NULL_T
        a.res = _NULL_T
        a.res2 = _NULL_T
        a.emode = a.res_from = 0
        a.y = S2m2Tensor(cv.data_ptr(), F32, 1, 1, w, w, w)
        a.w = f[B:].data_ptr()
        a.w_dtype = F32
        a.bias = None
        a.KH = a.KW = a.stride = 1
        a.pad_h = a.pad_w = 0
        a.Cout, a.act, a.shuffle2x2 = w, 0, 0
        a.batch, a.x_bs, a.w_bs, a.y_bs = B * h, w * Cc, w * Cc, w * w
        a.impl = 1
        self._chk(self.lib.s2m2b200_conv2d(C.byref(a), self._stream()), "row_corr")

    def row_corr_tc(self, f, gamma, beta, cv):
        """cv = all-pairs scan-line correlation of LayerNorm(f) (affine) on tcgen05 with split-fp16 operands (fp32-grade
        accuracy: the arg-max must not move); False if not applicable (caller: layernorm -> row_corr)."""
        if not self.row_corr_fused or f.dtype != torch.float16:
            return False
        B, h, w, _ = cv.shape
        nb = int(self.lib.s2m2b200_row_corr_workspace_bytes(B, h, w, f.shape[3]))
        if nb == 0 or not cv.is_contiguous() or cv.dtype != torch.float32:
            return False
        work = self._scratch(nb // 2, torch.float16, f.device)
        rc = self.lib.s2m2b200_row_corr(C.byref(tview(f)), C.c_void_p(gamma.data_ptr()), C.c_void_p(beta.data_ptr()), C.c_void_p(cv.data_ptr()),
                                        C.c_void_p(work.data_ptr()), C.c_int64(nb), self._stream())
        if rc == 1:
            return False
        self._chk(rc, "row_corr")
        return True

    def attention(self, qkv, o, S, N, heads, kv_shift, pe=None, pesum=None):
        q = tview(qkv)
        ot = tview(o)
        work, wbytes = None, 0
        if pesum is None and self.attn_tc:
            wbytes = int(self.lib.s2m2b200_attention_workspace_bytes(S, N, heads, o.shape[3] // heads, _DT[qkv.dtype]))
            if wbytes > 0:      # (the whole-row kernel needs no scratch any more: a cached token buffer keeps the ABI's opt-in)
                key = (str(qkv.device), wbytes)
                if key not in self._attn_work:
                    self._attn_work[key] = self._scratch(max(wbytes // 2, 8), torch.float16, qkv.device)
                work = C.c_void_p(self._attn_work[key].data_ptr())
        if pesum is not None:
            px, py, ph, pw = pe[:4]
            ind = pe[4] if len(pe) > 4 and self.attn_tc else None
            ps = S2m2Tensor(pesum.data_ptr(), _DT[pesum.dtype], 1, 1, S * N * heads, 32, 32)
            rc = self.lib.s2m2b200_attention(C.byref(q), C.byref(ot), S, N, heads, kv_shift, C.c_void_p(px.data_ptr()),
                                             C.c_void_p(py.data_ptr()), ph, pw, C.byref(ps),
                                             C.c_void_p(ind.data_ptr()) if ind is not None else None,
                                             C.c_int64(ind.numel() * 2 if ind is not None else 0), self._stream())
        else:
            rc = self.lib.s2m2b200_attention(C.byref(q), C.byref(ot), S, N, heads, kv_shift, None, None, 0, 0, None,
                                             work, C.c_int64(wbytes), self._stream())
        self._chk(rc, "attention")

    # -- norms / resampling / element-wise
    def layernorm(self, x, out, gamma=None, beta=None):
        g = C.c_void_p(gamma.data_ptr()) if gamma is not None else None
        b = C.c_void_p(beta.data_ptr()) if beta is not None else None
        xt, yt = tview(x), tview(out)
        self._chk(self.lib.s2m2b200_layernorm(C.byref(xt), C.byref(yt), g, b, self._stream()), "layernorm")

    def groupnorm(self, x, groups, gamma, beta, out):
        nblk = (x.shape[1] * x.shape[2] + 255) // 256
        key = (x.shape[0], groups, nblk, str(x.device))
        if key not in self._gn_work:
            self._gn_work[key] = self._scratch(2 * x.shape[0] * groups * (1 + nblk), torch.float64, x.device, zero=True)
        work = self._gn_work[key]
        xt, yt = tview(x), tview(out)
        self._chk(self.lib.s2m2b200_groupnorm(C.byref(xt), C.byref(yt), groups, C.c_void_p(gamma.data_ptr()),
                                              C.c_void_p(beta.data_ptr()), C.c_void_p(work.data_ptr()), self._stream()),
                  "groupnorm")

    def _two(self, fn, name, x, y, *extra):
        xt, yt = tview(x), tview(y)
        self._chk(fn(C.byref(xt), C.byref(yt), *extra, self._stream()), name)

    def avgpool2(self, x, out):
        self._two(self.lib.s2m2b200_avgpool2, "avgpool2", x, out)

    def upsample2x(self, x, out):
        self._two(self.lib.s2m2b200_upsample2x, "upsample2x", x, out)

    def copy_channels(self, src, dst):
        self._two(self.lib.s2m2b200_copy_channels, "copy_channels", src, dst)

    def unary(self, x, out, act):
        self._two(self.lib.s2m2b200_unary, "unary", x, out, C.c_int32(act))

    def gate_blend(self, f, g, z0, z1, out):
        ts = [tview(t) for t in (f, g, z0, z1, out)]
        self._chk(self.lib.s2m2b200_gate_blend(*[C.byref(t) for t in ts], self._stream()), "gate_blend")

    def mul(self, a, b, out):
        ts = [tview(t) for t in (a, b, out)]
        self._chk(self.lib.s2m2b200_mul(*[C.byref(t) for t in ts], self._stream()), "mul")

    def gru_update(self, z, q, h):
        ts = [tview(t) for t in (z, q, h)]
        self._chk(self.lib.s2m2b200_gru_update(*[C.byref(t) for t in ts], self._stream()), "gru_update")

    def preprocess(self, left, right, out):
        """left/right: (B,3,H,W) contiguous, float32 or uint8 (the same integer values either way)."""
        B, _, H, W = left.shape
        ot = tview(out)
        fn = self.lib.s2m2b200_preprocess_u8 if left.dtype == torch.uint8 else self.lib.s2m2b200_preprocess
        self._chk(fn(C.c_void_p(left.data_ptr()), C.c_void_p(right.data_ptr()), B, H, W, C.byref(ot), self._stream()), "preprocess")

    def stem(self, left, right, w0, b0, w1, b1, out):
        B, _, H, W = left.shape
        ot = tview(out)
        p = lambda t: C.c_void_p(t.data_ptr())
        fn = self.lib.s2m2b200_stem_u8 if left.dtype == torch.uint8 else self.lib.s2m2b200_stem
        self._chk(fn(p(left), p(right), B, H, W, p(w0), p(b0), p(w1), p(b1), C.byref(ot), self._stream()), "stem")

    def mask1x_inputs(self, up, rgb, wd, bd, wr, br, out):
        rt, ot = tview(rgb), tview(out)
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_mask1x_inputs(p(up), C.byref(rt), p(wd), p(bd), p(wr), p(br), C.byref(ot), self._stream()),
                  "mask1x_inputs")

    # -- matching
    def sinkhorn_softargmax(self, cv, disp, conf, occ, amax, use_positivity, iters):
        B, h, w, _ = cv.shape
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_sinkhorn_softargmax(p(cv), B, h, w, p(disp), p(conf), p(occ), p(amax),
                                                        int(use_positivity), iters, self._stream()), "sinkhorn_softargmax")

    def cv_lookup(self, cv, disp, out):
        B, h, w, _ = cv.shape
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_cv_lookup(p(cv), p(disp), B, h, w, p(out), self._stream()), "cv_lookup")

    # -- refiner glue
    def global_refiner_prep(self, disp, conf, out):
        ot = tview(out)
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_global_refiner_prep(p(disp), p(conf), C.c_int64(disp.numel()), C.byref(ot),
                                                        self._stream()), "global_refiner_prep")

    def global_refiner_merge(self, disp, conf, upd, use_positivity):
        if upd.dtype != torch.float32 or not upd.is_contiguous():
            raise ValueError("global_refiner_merge: update must be contiguous fp32")
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_global_refiner_merge(p(disp), p(conf), p(upd), C.c_int64(upd.shape[3]),
                                                         C.c_int64(disp.numel()), int(use_positivity), self._stream()),
                  "global_refiner_merge")

    def local_refiner_prep(self, disp, conf, occ, small):
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_local_refiner_prep(p(disp), p(conf), p(occ), C.c_int64(disp.numel()), p(small),
                                                       self._stream()), "local_refiner_prep")

    def local_refiner_update(self, disp, conf, occ, small, upd, use_positivity):
        B, h, w = disp.shape
        if upd.dtype != torch.float32 or not upd.is_contiguous() or upd.shape[3] < 3:
            raise ValueError("local_refiner_update: head output must be contiguous fp32 with >= 3 channels")
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_local_refiner_update(p(disp), p(conf), p(occ), p(small), p(upd), C.c_int64(upd.shape[3]),
                                                         B, h, w, int(use_positivity), self._stream()), "local_refiner_update")

    # -- upsampling
    def upsample4x_convex(self, disp, occ, conf, logits, up):
        B, h, w = disp.shape
        if logits.dtype != torch.float32 or not logits.is_contiguous() or logits.shape[3] < 9:
            raise ValueError("upsample4x_convex: logits must be contiguous fp32 (B,4h,4w,>=9)")
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_upsample4x_convex(p(disp), p(occ), p(conf), p(logits), logits.shape[3], B, h, w, p(up),
                                                      self._stream()), "upsample4x_convex")

    def filter1x(self, up, logits, out, upsample2=False):
        B, H, W, _ = up.shape
        if logits.dtype != torch.float32 or not logits.is_contiguous() or logits.shape[3] < 9:
            raise ValueError("filter1x: logits must be contiguous fp32 (B,H,W,>=9)")
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_filter1x(p(up), p(logits), logits.shape[3], B, H, W, p(out), int(upsample2), self._stream()),
                  "filter1x")

    def filter1x_crop(self, up, logits, out, crop_hw):
        """filter1x writing only the centre (crop_h, crop_w) window: image_crop (image_utils.py:66-96) fused into the last kernel."""
        B, H, W, _ = up.shape
        if logits.dtype != torch.float32 or not logits.is_contiguous() or logits.shape[3] < 9:
            raise ValueError("filter1x_crop: logits must be contiguous fp32 (B,H,W,>=9)")
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_filter1x_crop(p(up), p(logits), logits.shape[3], B, H, W, int(crop_hw[0]), int(crop_hw[1]), p(out),
                                                  self._stream()), "filter1x_crop")

    # -- pre-processing
    def image_pad(self, img, factor, out):
        """image_pad (image_utils.py:27-63) of a contiguous (B,3,H,W) uint8 / fp32 CUDA image into out (B,3,Hn,Wn) fp32."""
        B, _, H, W = img.shape
        work = torch.empty((B * 3 * (H // factor) * (W // factor),), dtype=torch.float32, device=img.device)
        self._chk(self.lib.s2m2b200_image_pad(C.c_void_p(img.data_ptr()), 1 if img.dtype == torch.uint8 else 0, B, H, W, int(factor),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(work.data_ptr()), self._stream()), "image_pad")

    def remap_u8(self, src, mapx, mapy, out, planar=True):
        """cv2.remap(uint8 HWC image, INTER_LINEAR, border 0) on the device, bit-exact (image_utils.py:108-136).  src: (Hs,Ws,3) shared
        by all N maps, or (N,Hs,Ws,3); mapx / mapy: (N,Hd,Wd) fp32; out: (N,3,Hd,Wd) planar or (N,Hd,Wd,3) uint8."""
        N, Hd, Wd = mapx.shape
        Hs, Ws = src.shape[-3], src.shape[-2]
        stride = Hs * Ws * 3 if src.dim() == 4 else 0
        p = lambda t: C.c_void_p(t.data_ptr())
        self._chk(self.lib.s2m2b200_remap_u8(p(src), C.c_int64(stride), Hs, Ws, p(mapx), p(mapy), N, Hd, Wd, p(out), int(planar), self._stream()),
                  "remap_u8")

    # -- consumers of the output maps
    def confidence_score(self, conf, crop_hw, margin, out):
        """out[b] = mean of conf[b] over the centre crop `crop_hw` shrunk by `margin` (model_utils.py:93-94)."""
        B, H, W = conf.shape[0], conf.shape[-2], conf.shape[-1]
        self._chk(self.lib.s2m2b200_confidence_score(C.c_void_p(conf.data_ptr()), B, H, W, int(crop_hw[0]), int(crop_hw[1]), int(margin),
                                                     C.c_void_p(out.data_ptr()), self._stream()), "confidence_score")

    def disparity_to_points(self, disp, conf, occ, fx, cx, cy, baseline, doffs, depth_scale, depth_trunc, xyz, valid):
        B, H, W = disp.shape[0], disp.shape[-2], disp.shape[-1]
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        f = C.c_float
        self._chk(self.lib.s2m2b200_disparity_to_points(p(disp), p(conf), p(occ), B, H, W, f(fx), f(cx), f(cy), f(baseline), f(doffs),
                                                        f(depth_scale), f(depth_trunc), p(xyz), p(valid), self._stream()),
                  "disparity_to_points")
