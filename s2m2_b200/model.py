"""``S2M2`` — drop-in replacement for the reference's ``s2m2.core.model.s2m2.S2M2`` nn.Module.

Same constructor, ``my_load_state_dict``, ``eval``/``to`` and ``forward(img0, img1) ->
(disparity, occlusion, confidence)`` surface (reference: s2m2.py:14-78,136-197; callers:
model_utils.py:30-43,72-83), same parameter names/shapes — but ``forward`` is a schedule of
hand-written sm_100a CUDA kernels reached through the C-ABI in ``include/s2m2b200.h``
(``s2m2_b200.ops.CudaOps``).  There is no PyTorch-op or CPU fallback: without the built library and
a CUDA device ``forward`` raises.

Everything is NHWC on the device.  ``precision``:
  * ``"fp32"``  — fp32 activations/weights, fp32 FMA GEMMs (parity mode, the oracle's arithmetic);
  * ``"fp16"``  — fp16 activations/weights, tensor-core GEMMs with fp32 accumulation; the matching
    maps (cost volume, Sinkhorn, disparity/confidence/occlusion, upsampling) stay fp32.

The op backend is injectable (``model.ops``) so that tests can run this exact schedule against a
plain-torch restatement of every op (``tests/torch_ops.py``) on CPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .weights import param_spec

ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4


class ConvW:
    """Packed weights of one implicit-GEMM convolution: ``w`` is (Cout, KH*KW*Cin) K-major, tap-major."""
    __slots__ = ("w", "b", "kh", "kw", "stride", "ph", "pw", "cin", "cout", "shuffle", "name")

    def __init__(self, w, b, kh, kw, stride, ph, pw, cin, cout, shuffle=False, name=""):
        self.w, self.b, self.kh, self.kw, self.stride, self.ph, self.pw = w, b, kh, kw, stride, ph, pw
        self.cin, self.cout, self.shuffle, self.name = cin, cout, shuffle, name


class S2M2(nn.Module):
    def __init__(self, feature_channels: int, dim_expansion: int = 1, num_transformer: int = 1,
                 use_positivity: bool = False, output_upsample: bool = False, refine_iter: int = 3,
                 precision: str = "fp16"):
        super().__init__()
        if dim_expansion != 1:
            raise ValueError("only dim_expansion=1 is supported (the reference's load_model always passes 1)")
        if precision not in ("fp32", "fp16"):
            raise ValueError("precision must be 'fp32' or 'fp16'")
        self.feature_channels = feature_channels
        self.num_transformer = num_transformer
        self.use_positivity = use_positivity
        self.refine_iter = refine_iter
        self.output_upsample = output_upsample
        self.precision = precision
        # flat parameter registry under the reference's names ('.' is not allowed in parameter names,
        # so the reference keys are kept in a side table and exposed through state_dict()).
        self._spec = param_spec(feature_channels, num_transformer)
        self._p = nn.ParameterDict()
        for name, shape in self._spec:
            self._p[name.replace(".", "/")] = nn.Parameter(torch.zeros(shape), requires_grad=False)
        self.ops = None            # op backend (CudaOps); created lazily on first forward
        self._packed = None        # packed weights (device, precision) — rebuilt when parameters move/change
        self._packed_key = None
        self.taps: Optional[dict] = None   # set to a dict to collect stage outputs (tests)
        import os
        self.auto_graph = os.environ.get("S2M2_AUTO_GRAPH", "1") != "0"   # replay a captured CUDA graph on repeated same-shape calls
        self.auto_graph_slots = int(os.environ.get("S2M2_AUTO_GRAPH_SLOTS", "2"))
        self.auto_graph_warm = int(os.environ.get("S2M2_AUTO_GRAPH_WARM", "1"))   # eager calls of a shape before its graph is captured
        self._graphs = {}

    # ---- nn.Module surface -------------------------------------------------------------------
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """nn.Module.state_dict under the REFERENCE key names (the parameters live in a flat ParameterDict because '.' is not
        allowed in parameter names); honours ``destination`` / ``prefix`` / ``keep_vars`` like torch."""
        if args:   # legacy positional form (destination, prefix, keep_vars)
            destination = args[0] if len(args) > 0 else destination
            prefix = args[1] if len(args) > 1 else prefix
            keep_vars = args[2] if len(args) > 2 else keep_vars
        from collections import OrderedDict
        out = OrderedDict() if destination is None else destination
        for name, _ in self._spec:
            p = self._p[name.replace(".", "/")]
            out[prefix + name] = p if keep_vars else p.detach()
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        own = {name for name, _ in self._spec}
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"state_dict mismatch: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        for name, shape in self._spec:
            if name in state_dict:
                t = state_dict[name]
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {name}: {tuple(t.shape)} vs {tuple(shape)}")
                p = self._p[name.replace(".", "/")]
                p.data = t.detach().to(device=p.device, dtype=torch.float32).clone()
        self._packed = None
        self._graphs = {}
        from torch.nn.modules.module import _IncompatibleKeys
        return _IncompatibleKeys(missing, unexpected)

    def my_load_state_dict(self, state_dict):
        """Tolerant load, same behaviour as s2m2.py:69-78: shape-mismatched keys are skipped with a print."""
        own = dict(self._spec)
        sd = dict(state_dict)
        for k in list(sd):
            if k in own and tuple(sd[k].shape) != tuple(own[k]):
                print(f"Skip loading parameter: {k}, required shape: {torch.Size(own[k])}, "
                      f"loaded shape: {sd[k].shape}")
                del sd[k]
        self.load_state_dict(sd, strict=False)

    def _apply(self, fn, recurse=True):
        self._packed = None
        self._graphs = {}
        return super()._apply(fn, recurse)

    # ---- weight packing ------------------------------------------------------------------------
    def _w(self, name):
        return self._p[name.replace(".", "/")].data

    def _has(self, name):
        return name.replace(".", "/") in self._p

    def _adt(self):
        return torch.float16 if self.precision == "fp16" else torch.float32

    def _mk(self, w4, b, stride=1, ph=None, pw=None, shuffle=False, name="", cin_to=None, cout_to=None):
        """w4: (Cout, Cin, KH, KW) fp32 conv weight -> ConvW (K order = tap-major, channel-minor).

        ``cin_to`` / ``cout_to`` zero-pad the channel counts (to multiples of 64 so the layer runs on the tcgen05
        kernel); padded outputs are act(0 + 0) = 0 for GELU/ReLU and meet zero weights downstream."""
        if cin_to is not None and cin_to > w4.shape[1]:
            w4 = torch.cat([w4, w4.new_zeros(w4.shape[0], cin_to - w4.shape[1], *w4.shape[2:])], 1)
        if cout_to is not None and cout_to > w4.shape[0]:
            pad = cout_to - w4.shape[0]
            w4 = torch.cat([w4, w4.new_zeros(pad, *w4.shape[1:])], 0)
            b = torch.cat([b, b.new_zeros(pad)]) if b is not None else None
        cout, cin, kh, kw = w4.shape
        ph = kh // 2 if ph is None else ph
        pw = kw // 2 if pw is None else pw
        wp = w4.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin).contiguous().to(self._adt())
        bb = None if b is None else b.detach().float().contiguous()
        return ConvW(wp, bb, kh, kw, stride, ph, pw, cin, cout, shuffle, name)

    def _pack_d0(self, name):
        """conv1_down.0 = Conv2d(16, 64, 5, stride 2, pad 2) (submodules.py:74).  fp32: as is.  fp16: the stem writes its 16
        channels in space-to-depth layout (2x2 pixel blocks -> 64 channels at half resolution), which turns this layer into
        a dense 3x3 stride-1 pad-1 convolution over 64 channels (K = 576) for the tcgen05 kernel:
        W'[co, (py*2+px)*16 + c, ty, tx] = W[co, c, 2*ty + py, 2*tx + px]   (taps with 2*t + p == 5 are zero)."""
        w, b = self._w(name + ".weight"), self._w(name + ".bias")
        if self.precision != "fp16":
            return self._mk(w, b, stride=2, name=name)
        w3 = w.new_zeros(w.shape[0], 64, 3, 3)
        for ty in range(3):
            for py in range(2):
                for tx in range(3):
                    for px in range(2):
                        ky, kx = 2 * ty + py, 2 * tx + px
                        if ky < 5 and kx < 5:
                            c0 = (py * 2 + px) * 16
                            w3[:, c0:c0 + 16, ty, tx] = w[:, :, ky, kx]
        return self._mk(w3, b, stride=1, name=name)

    def _conv(self, name, stride=1, **kw):
        b = self._w(name + ".bias") if self._has(name + ".bias") else None
        return self._mk(self._w(name + ".weight"), b, stride=stride, name=name, **kw)

    def _convT1(self, name, **kw):
        w = self._w(name + ".weight")
        return self._mk(w.permute(1, 0, 2, 3).contiguous(), self._w(name + ".bias"), name=name, **kw)

    def _lin(self, name):
        w = self._w(name + ".weight")
        b = self._w(name + ".bias") if self._has(name + ".bias") else None
        return self._mk(w[:, :, None, None], b, name=name)

    def _cat_out(self, convs, name):
        """Fuse convs sharing one input into a single GEMM (concatenate along Cout)."""
        w = torch.cat([c.w for c in convs], 0)
        dev = w.device
        bs = [c.b if c.b is not None else torch.zeros(c.cout, device=dev) for c in convs]
        c0 = convs[0]
        return ConvW(w.contiguous(), torch.cat(bs).contiguous(), c0.kh, c0.kw, c0.stride, c0.ph, c0.pw, c0.cin,
                     sum(c.cout for c in convs), False, name)

    def _convT2(self, name, cout_to=None):
        """ConvTranspose2d(k=2,s=2): weight (Cin,Cout,2,2) -> GEMM (4*Cout, Cin) + pixel-shuffle store.
        ``cout_to`` zero-pads the output channels (9 -> 16 logits so the layer takes the tcgen05 path)."""
        w = self._w(name + ".weight")
        b = self._w(name + ".bias")
        if cout_to is not None and cout_to > w.shape[1]:
            pad = cout_to - w.shape[1]
            w = torch.cat([w, w.new_zeros(w.shape[0], pad, 2, 2)], 1)
            b = torch.cat([b, b.new_zeros(pad)])
        cin, cout = w.shape[:2]
        wg = w.permute(2, 3, 1, 0).reshape(4 * cout, cin)          # row = (a*2+b)*Cout + co
        c = self._mk(wg[:, :, None, None], b.repeat(4), shuffle=True, name=name)
        return c

    def _convT3(self, name):
        """ConvTranspose2d(k=3,p=1,s=1) == 3x3 conv with flipped, in/out-transposed weights."""
        w = self._w(name + ".weight")
        return self._mk(w.flip(2, 3).permute(1, 0, 2, 3).contiguous(), self._w(name + ".bias"), name=name)

    def _head_pad(self):
        """Tiny-Cout heads (C -> 1/2/3, K = 9C) are GEMV-like: on the tcgen05 path they run as one 64-wide tile
        (zero weights for the padding); in fp32 mode the CUDA-core kernel takes them unpadded."""  # tile is 64 wide, 16 columns are stored
        return 16 if self.precision == "fp16" else None

    def _pack_update_heads(self, r, C):
        """disp_update.2 (C->1) and conf_occ_update.2 (C->2), both bias-free 3x3 (refinenet.py:115-122), as ONE conv over
        the fused hidden tensor [disp hidden (C) | conf/occ hidden (C)] with block-structured weights -> 3 channels."""
        wd = self._w(f"{r}.disp_update.2.weight")          # (1, C, 3, 3)
        wc = self._w(f"{r}.conf_occ_update.2.weight")      # (2, C, 3, 3)
        w = torch.zeros(3, 2 * C, 3, 3, device=wd.device)
        w[0:1, :C] = wd
        w[1:3, C:] = wc
        return self._mk(w, None, name=f"{r}.update_heads", cout_to=self._head_pad())

    def _pack_attn(self, p, heads, pe=False):
        q, k, v = self._lin(p + ".q"), self._lin(p + ".k"), self._lin(p + ".v")
        d = {"qkv": self._cat_out([q, k, v], p + ".qkv"), "proj": self._lin(p + ".proj"), "heads": heads}
        if pe:
            d["pe_proj"] = self._lin(p + ".pe_proj")
        return d

    def _pack_ffn(self, p):
        return {"l0": self._lin(p + ".ffn.0"), "l2": self._lin(p + ".ffn.2")}

    def _pack_basic(self, p, heads):
        return {"cross": self._pack_attn(p + ".cross_attn.attn", heads), "ffn_c": self._pack_ffn(p + ".ffn_c"),
                "self": self._pack_attn(p + ".self_attn.attn", heads), "ffn": self._pack_ffn(p + ".ffn")}

    def _pack_global(self, p, heads, cross, pe):
        d = {"self": self._pack_attn(p + ".self_attn.attn", heads, pe), "ffn": self._pack_ffn(p + ".ffn")}
        if cross:
            d["cross"] = self._pack_attn(p + ".cross_attn.attn", heads)
            d["ffn_c"] = self._pack_ffn(p + ".ffn_c")
        return d

    def _pack_convblock(self, p):
        return {"c0": self._conv(p + ".convs.0"), "c2": self._conv(p + ".convs.2"),
                "x0": self._conv(p + ".convs_1x.0"), "x2": self._conv(p + ".convs_1x.2")}

    def _pack_fusion(self, p):
        g0, f0 = self._conv(p + ".feature_gate.0"), self._conv(p + ".feature_fusion.0")
        return {"gf0": self._cat_out([g0, f0], p + ".gf0"), "g2": self._conv(p + ".feature_gate.2"),
                "f2": self._conv(p + ".feature_fusion.2"), "dim": g0.cout}

    def _pack_updown(self, p):
        return {f"{n}{i}": self._conv(f"{p}.{n}_conv{i}.1") for n in ("down", "up") for i in range(3)}

    def _pack_unet(self, p, n_attn, pe):
        d = self._pack_updown(p)
        for i in range(3):
            d[f"cat{i}"] = self._pack_fusion(f"{p}.concat_conv{i}")
            d[f"enc{i}"] = self._pack_convblock(f"{p}.enc{i}")
            d[f"dec{i}"] = self._pack_convblock(f"{p}.dec{i}")
        d["enc3s"] = [self._pack_global(f"{p}.enc3s.{i}", 8, False, pe) for i in range(n_attn)]
        d["dec3s"] = [self._pack_global(f"{p}.dec3s.{i}", 8, False, False) for i in range(n_attn)]
        return d

    def _pack_mrt(self, p):
        d = self._pack_updown(p)
        for i in (1, 2, 3):
            d[f"dcat{i}"] = self._pack_fusion(f"{p}.down_concat{i}")
        for i in (0, 1, 2):
            d[f"ucat{i}"] = self._pack_fusion(f"{p}.up_concat{i}")
            d[f"enc{i}"] = self._pack_basic(f"{p}.enc_attn{i}", 1 << i)
            d[f"dec{i}"] = self._pack_basic(f"{p}.dec_attn{i}", 1 << i)
        d["enc3s"] = [self._pack_global(f"{p}.enc_attn3s.{i}", 8, True, False) for i in range(2)]
        d["dec3s"] = [self._pack_global(f"{p}.dec_attn3s.{i}", 8, True, False) for i in range(2)]
        return d

    def _pack(self):
        C, NTR = self.feature_channels, self.num_transformer
        P: Dict[str, object] = {}
        b = "cnn_backbone"
        f32c = lambda n: self._w(n).float().contiguous()
        hostc = lambda t: t.detach().float().cpu().contiguous()      # tiny weights that travel as kernel parameters
        P["bb"] = {"w0": hostc(self._w(f"{b}.conv0.0.weight").reshape(16, 3)), "b0": hostc(self._w(f"{b}.conv0.0.bias")),
                   "w1": hostc(self._w(f"{b}.conv0.2.weight").reshape(16, 16)), "b1": hostc(self._w(f"{b}.conv0.2.bias")),
                   "d0": self._pack_d0(f"{b}.conv1_down.0"),
                   "d2": self._conv(f"{b}.conv1_down.2"),
                   "gn_w": self._w(f"{b}.norm1.weight").float().contiguous(),
                   "gn_b": self._w(f"{b}.norm1.bias").float().contiguous(),
                   "c20": self._conv(f"{b}.conv2.0"), "c22": self._conv(f"{b}.conv2.2"),
                   "down": self._conv(f"{b}.conv2_down.0", stride=2)}
        P["pyr"] = self._pack_unet("feat_pyramid", 2 * NTR, True)
        P["mrt"] = [self._pack_mrt(f"transformer.uformer_list.{l}") for l in range(NTR)]
        P["di_w"] = self._w("disp_init.layer_norm.weight").float().contiguous()
        P["di_b"] = self._w("disp_init.layer_norm.bias").float().contiguous()
        # global refiner: input channel order in the reference is [disp, conf_logit, ctx(C)] (refinenet.py:68);
        # we run the conv over two sources [ctx(C) | aux(64): disp, conf_logit, 62 zeros]: ctx is read in place
        # (no concat copy) and both sources are multiples of 64 channels (tcgen05 path).
        g = "global_refiner"
        w = self._w(f"{g}.init_feat.0.weight")                        # (C, 2+C, 3, 3)
        wpad = torch.zeros(w.shape[0], C + 64, 3, 3, device=w.device)
        wpad[:, :C] = w[:, 2:]
        wpad[:, C:C + 2] = w[:, :2]
        out_w = self._w(f"{g}.out_feat.0.weight") * 1e2               # fold the *1e2 (refinenet.py:71)
        out_b = self._w(f"{g}.out_feat.0.bias") * 1e2
        P["gr"] = {"i0": self._mk(wpad, self._w(f"{g}.init_feat.0.bias"), name=f"{g}.init_feat.0"),
                   "i2": self._conv(f"{g}.init_feat.2"), "unet": self._pack_unet(f"{g}.refine_unet", 1, False),
                   "out": self._mk(out_w, out_b, name=f"{g}.out_feat.0", cout_to=self._head_pad())}
        P["ffl"] = self._pack_fusion("feat_fusion_layer")
        P["ctx0"], P["ctx2"] = self._conv("ctx_feat.0"), self._conv("ctx_feat.2")
        r = "refiner"
        P["lr"] = {"df0": self._conv(f"{r}.disp_feat.0", cout_to=128), "df2": self._conv(f"{r}.disp_feat.2", cin_to=128),
                   "cf10": self._conv(f"{r}.corr_feat1.0", cout_to=128), "cf12": self._conv(f"{r}.corr_feat1.2", cin_to=128),
                   "cf20": self._conv(f"{r}.corr_feat2.0", cout_to=128), "cf22": self._conv(f"{r}.corr_feat2.2", cin_to=128),
                   "co0": self._conv(f"{r}.conf_occ_feat.0"), "co2": self._conv(f"{r}.conf_occ_feat.2"),
                   "cat0": self._conv(f"{r}.disp_corr_ctx_cat.0"), "cat2": self._conv(f"{r}.disp_corr_ctx_cat.2"),
                   "unet": self._pack_unet(f"{r}.refine_unet", 1, False),
                   "upd0": self._cat_out([self._conv(f"{r}.disp_update.0"), self._conv(f"{r}.conf_occ_update.0")],
                                         f"{r}.upd0"),
                   "upd2": self._pack_update_heads(r, C),
                   "zr1": self._cat_out([self._conv(f"{r}.gru.convz1", ph=1, pw=0), self._conv(f"{r}.gru.convr1", ph=1, pw=0)], "zr1"),
                   "q1": self._conv(f"{r}.gru.convq1", ph=1, pw=0),
                   "zr2": self._cat_out([self._conv(f"{r}.gru.convz2", ph=0, pw=1), self._conv(f"{r}.gru.convr2", ph=0, pw=1)], "zr2"),
                   "q2": self._conv(f"{r}.gru.convq2", ph=0, pw=1)}
        u = "upsample_mask_4x_refine"
        P["u4"] = {"x": self._convT2(f"{u}.conv_x"), "y": self._conv(f"{u}.conv_y"),
                   "c0": self._conv(f"{u}.conv_concat.0"), "c2": self._convT2(f"{u}.conv_concat.2", cout_to=16)}
        u = "upsample_mask_1x"
        wd = self._w(f"{u}.conv_disp.0.weight").flip(2, 3).permute(1, 2, 3, 0).reshape(16, 9)      # [co][tap]
        wr = self._w(f"{u}.conv_rgb.0.weight").flip(2, 3).permute(1, 2, 3, 0).reshape(16, 27)     # [co][tap*3+ci]
        P["u1"] = {"wd": hostc(wd), "bd": hostc(self._w(f"{u}.conv_disp.0.bias")),
                   "wr": hostc(wr), "br": hostc(self._w(f"{u}.conv_rgb.0.bias")),
                   "ctx": self._convT2(f"{u}.conv_ctx"), "c0": self._conv(f"{u}.conv_concat.0", cin_to=64, cout_to=64),
                   "c2": self._convT1(f"{u}.conv_concat.2", cin_to=64, cout_to=16)}
        return P

    def _ensure_packed(self, device):
        key = (str(device), self.precision, self.use_positivity)
        if self._packed is None or self._packed_key != key:
            if self._w("ctx_feat.0.weight").device != device:
                raise RuntimeError(f"model parameters are on {self._w('ctx_feat.0.weight').device}, inputs on {device}")
            self._packed = self._pack()
            self._packed_key = key

    # ---- building blocks (NHWC) ----------------------------------------------------------------
    def _new(self, ref, C, dtype=None, hw=None):
        n, h, w = ref.shape[0], ref.shape[1], ref.shape[2]
        if hw is not None:
            h, w = hw
        return self.ops.empty((n, h, w, C), dtype or self._adt(), ref.device)

    def _c(self, x, cw: ConvW, act=ACT_NONE, out=None, res=None, x2=None, out_dtype=None):
        """Run one convolution.  ``x2``: second input whose channels follow x's (conv over cat([x,x2]))."""
        if out is None:
            if cw.shuffle:
                out = self._new(x, cw.cout // 4, out_dtype, (2 * x.shape[1], 2 * x.shape[2]))
            else:
                ho = (x.shape[1] + 2 * cw.ph - cw.kh) // cw.stride + 1
                wo = (x.shape[2] + 2 * cw.pw - cw.kw) // cw.stride + 1
                out = self._new(x, cw.cout, out_dtype, (ho, wo))
        self.ops.conv2d(x, cw, act, out, res, x2)
        return out

    def _ln(self, x, gamma=None, beta=None, out_dtype=None):
        out = self._new(x, x.shape[3], out_dtype)
        self.ops.layernorm(x, out, gamma, beta)
        return out

    def _ffn(self, z, p):
        if self.ops.ffn(z, p["l0"], p["l2"]):        # fused tcgen05 kernel (fp16): LN, both GEMMs, GELU and the residual
            return z
        h = self._c(self._ln(z), p["l0"], ACT_GELU)
        self._c(h, p["l2"], out=z, res=z)            # in place: z += ffn(LN(z))   (attentions.py:245-250)
        return z

    def _attn(self, z, p, S, N, cross, out=None, pe=None):
        """LN -> fused QKV GEMM -> attention -> proj + residual (attentions.py:33-54,77-96,119-161)."""
        C = z.shape[3]
        heads = p["heads"]
        qkv = self._new(z, p["qkv"].cout)
        if not self.ops.ln_linear(z, p["qkv"], qkv):   # fused tcgen05 kernel (fp16, >= 16384 tokens): LN(z) stays in smem
            self._c(self._ln(z), p["qkv"], out=qkv)
        o = self._new(z, C)
        pesum = None
        if pe is not None and "pe_proj" in p:
            pesum = self.ops.empty((S, N, heads, 32), self._adt(), z.device)
        self.ops.attention(qkv, o, S, N, heads, S // 2 if cross else 0, pe, pesum)
        if pesum is not None:
            d = C // heads
            # out_head += pe_proj(pe_sum): one GEMM over (tokens*heads, 32) -> (tokens*heads, d)  (attentions.py:47-48)
            ov = o.reshape(1, 1, S * N * heads, d)
            self._c(pesum.reshape(1, 1, S * N * heads, 32), p["pe_proj"], out=ov, res=ov)
        out = z if out is None else out
        self._c(o, p["proj"], out=out, res=z)
        return out

    def _basic_block(self, z, p, fresh):
        """BasicAttnBlock (attentions.py:347-355) on NHWC z (2B,h,w,C); 1-D attention along w."""
        n, h, w, C = z.shape
        out = self._new(z, C) if fresh else z
        z = self._attn(z, p["cross"], n * h, w, True, out=out)
        z = self._ffn(z, p["ffn_c"])
        z = self._attn(z, p["self"], n * h, w, False)
        return self._ffn(z, p["ffn"])

    def _global_block(self, z, p, fresh, pe=None):
        """GlobalAttnBlock (attentions.py:311-321) on NHWC z; 2-D attention over all h*w tokens."""
        n, h, w, C = z.shape
        out = self._new(z, C) if fresh else z
        if "cross" in p:
            z = self._attn(z, p["cross"], n, h * w, True, out=out)
            z = self._ffn(z, p["ffn_c"])
            z = self._attn(z, p["self"], n, h * w, False, pe=pe)
        else:
            z = self._attn(z, p["self"], n, h * w, False, out=out, pe=pe)
        return self._ffn(z, p["ffn"])

    def _convblock(self, z, p, out=None):
        """ConvBlock2D (attentions.py:277-281)."""
        a = self._c(self._c(z, p["c0"], ACT_GELU), p["c2"])
        o = a if out is None else out
        if self.ops.mlp2(z, a, o, p["x0"], p["x2"], ACT_RELU):     # fused 1x1 -> ReLU -> 1x1 + a (hidden tile stays in smem)
            return o
        u = self._c(z, p["x0"], ACT_RELU)
        return self._c(u, p["x2"], out=out, res=a)

    def _fusion(self, z0, z1, p):
        """FeatureFusion (feature_fusion.py:24-33); conv over cat([z0,z1]) reads both sources directly."""
        d = p["dim"]
        t = self._c(z0, p["gf0"], ACT_GELU, x2=z1)                   # (.., 3d): gate hidden | fusion hidden
        if self.ops.conv_emodes(t):
            # the gate's conv epilogue forms  clamp(sigmoid(.)) z0 + (1 - clamp) z1  (mode 2); the fusion conv adds it as its residual
            bl = self._new(z0, d)
            if self.ops.conv2d(t[..., :d], p["g2"], ACT_SIGMOID, bl, res=z0, res2=z1, emode=self.ops.EMODE_BLEND):
                return self._c(t[..., d:], p["f2"], res=bl)
        g = self._c(t[..., :d], p["g2"], ACT_SIGMOID)
        f = self._c(t[..., d:], p["f2"])
        self.ops.gate_blend(f, g, z0, z1, f)
        return f

    def _down(self, z, cw):
        """AvgPool2d(2) -> 1x1 conv (unet.py:25-30)."""
        n, h, w, C = z.shape
        p = self._new(z, C, hw=(h // 2, w // 2))
        self.ops.avgpool2(z, p)
        return self._c(p, cw)

    def _up(self, z, cw):
        """bilinear x2 -> 1x1 conv (unet.py:32-37), evaluated as 1x1 conv -> bilinear x2 (both linear, the
        interpolation weights sum to 1 so the bias commutes): 4x fewer GEMM FLOPs."""
        t = self._c(z, cw)
        n, h, w, C = t.shape
        u = self._new(t, C, hw=(2 * h, 2 * w))
        self.ops.upsample2x(t, u)
        return u

    def _unet(self, z, p, pe=None, tap=None, out0=None):
        """Unet.forward (unet.py:65-112).  ``out0``: optional destination view of the 1/4-resolution decoder output."""
        z0 = self._convblock(z, p["enc0"])
        z1 = self._convblock(self._down(z0, p["down0"]), p["enc1"])
        z2 = self._convblock(self._down(z1, p["down1"]), p["enc2"])
        z3 = self._down(z2, p["down2"])
        for i, bp in enumerate(p["enc3s"]):
            z3 = self._global_block(z3, bp, False, pe)
            if tap is not None and self.taps is not None:
                self.taps[f"{tap}.enc3s.{i}"] = z3.clone()
        for bp in p["dec3s"]:
            z3 = self._global_block(z3, bp, False)
        z2n = self._convblock(self._fusion(z2, self._up(z3, p["up2"]), p["cat2"]), p["dec2"])
        z1n = self._convblock(self._fusion(z1, self._up(z2n, p["up1"]), p["cat1"]), p["dec1"])
        z0n = self._convblock(self._fusion(z0, self._up(z1n, p["up0"]), p["cat0"]), p["dec0"], out=out0)
        return z0n, z1n, z2n, z3

    def _mrt(self, z0, z1, z2, z3, p):
        """MRT.forward (stacked_MRT.py:89-133)."""
        z0 = self._basic_block(z0, p["enc0"], True)
        z1 = self._basic_block(self._fusion(z1, self._down(z0, p["down0"]), p["dcat1"]), p["enc1"], False)
        z2 = self._basic_block(self._fusion(z2, self._down(z1, p["down1"]), p["dcat2"]), p["enc2"], False)
        z3 = self._fusion(z3, self._down(z2, p["down2"]), p["dcat3"])
        for bp in p["enc3s"] + p["dec3s"]:
            z3 = self._global_block(z3, bp, False)
        z2 = self._basic_block(self._fusion(z2, self._up(z3, p["up2"]), p["ucat2"]), p["dec2"], False)
        z1 = self._basic_block(self._fusion(z1, self._up(z2, p["up1"]), p["ucat1"]), p["dec1"], False)
        z0 = self._basic_block(self._fusion(z0, self._up(z1, p["up0"]), p["ucat0"]), p["dec0"], False)
        return z0, z1, z2, z3

    def _tap(self, key, t, nhwc=True):
        if self.taps is not None:
            self.taps[key] = t.float().permute(0, 3, 1, 2).contiguous() if nhwc else t.float().clone()

    # ---- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, img0: torch.Tensor, img1: torch.Tensor, crop_hw=None):
        """(B,3,H,W) float [0,255] x2 -> disparity, occlusion, confidence, each (B,1,H,W) fp32 (s2m2.py:136-197).
        ``crop_hw`` (not in the reference's signature; default None = the reference's behaviour): the un-padded image size —
        the last kernel then writes only that centre window, i.e. image_crop (image_utils.py:66-96, the step
        run_stereo_matching applies to each output, model_utils.py:89-91) is fused into it.

        Repeated calls with the same (shape, dtype, crop) replay a CUDA graph captured at the end of the first such call (the caller's
        warm-up call): the ~900 kernel launches of one forward then cost one graph launch on the host — what the reference's caller
        (run_stereo_matching's timing loop, model_utils.py:79-83; the calibration optimisers' ~100 calls, cem.py:55-72) sees
        through the unchanged nn.Module surface.  ``model.auto_graph = False`` (or S2M2_AUTO_GRAPH=0) keeps every call eager."""
        if (self.auto_graph and self.taps is None and img0.is_cuda and self.ops is not None and hasattr(self.ops, "lib")
                and not torch.cuda.is_current_stream_capturing()):
            return self._forward_graphed(img0, img1, crop_hw)
        return self._forward_eager(img0, img1, crop_hw)

    def _forward_graphed(self, img0, img1, crop_hw):
        key = (tuple(img0.shape), img0.dtype, img1.dtype, str(img0.device), None if crop_hw is None else tuple(int(v) for v in crop_hw),
               self.precision, self.use_positivity, self.refine_iter, self.output_upsample)
        ent = self._graphs.get(key)
        if ent is not None and "graph" in ent:
            self._graphs[key] = self._graphs.pop(key)                  # most recently used last
            ent["in0"].copy_(img0); ent["in1"].copy_(img1)
            ent["graph"].replay()
            return tuple(t.clone() for t in ent["out"])               # the graph's output buffers are overwritten by the next replay
        if ent is None:
            ent = self._graphs[key] = {"n": 0}
            while len(self._graphs) > self.auto_graph_slots:           # a captured graph pins its whole activation pool: keep few
                self._graphs.pop(next(iter(self._graphs)))
        res = self._forward_eager(img0, img1, crop_hw)                # (the first call also packs the weights / builds the tables)
        ent["n"] += 1
        if ent["n"] >= self.auto_graph_warm and tuple(img0.shape) == tuple(img1.shape):
            # capture now, at the end of the warm-up call(s): the caller's NEXT call (run_stereo_matching's timed loop after its
            # warm-up call, model_utils.py:69-70 of the demo) is already a replay
            dev = img0.device
            ent["in0"] = torch.empty_like(img0, memory_format=torch.contiguous_format)
            ent["in1"] = torch.empty_like(img1, memory_format=torch.contiguous_format)
            ent["in0"].copy_(img0); ent["in1"].copy_(img1)
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    out = self._forward_eager(ent["in0"], ent["in1"], crop_hw)
            cur.wait_stream(side)
            ent["graph"], ent["out"] = g, out
        return res

    def _forward_eager(self, img0: torch.Tensor, img1: torch.Tensor, crop_hw=None):
        if self.ops is None:
            from .ops import CudaOps
            self.ops = CudaOps()                      # raises if the library / GPU is missing: no fallback
        if img0.shape != img1.shape or img0.dim() != 4 or img0.shape[1] != 3:
            raise ValueError("expected two (B,3,H,W) images of equal shape")
        B, _, H, W = img0.shape
        if H % 32 or W % 32:
            raise ValueError("H and W must be multiples of 32 (run_stereo_matching pads, model_utils.py:72-73)")
        dev = img0.device
        self._ensure_packed(dev)
        P, ops, C, adt = self._packed, self.ops, self.feature_channels, self._adt()
        f32 = torch.float32
        h4, w4 = H // 4, W // 4

        # -- normalise + NCHW->NHWC; left and right stacked on the batch dim (s2m2.py:86-96)
        img = ops.empty((2 * B, H, W, 3), adt, dev)                      # normalised NHWC copy (mask head input)
        # float32 (what image_pad produces) or uint8 images (4x fewer input bytes, bit-identical results)
        u8 = img0.dtype == torch.uint8 and img1.dtype == torch.uint8
        in0, in1 = (img0.contiguous(), img1.contiguous()) if u8 else (img0.float().contiguous(), img1.float().contiguous())
        ops.preprocess(in0, in1, img)

        # -- CNN backbone (submodules.py:88-93)
        bb = P["bb"]
        # fp16: space-to-depth stem output (2B,H/2,W/2,64), see _pack_d0; fp32: (2B,H,W,16)
        x = ops.empty((2 * B, H // 2, W // 2, 64) if self.precision == "fp16" else (2 * B, H, W, 16), adt, dev)
        ops.stem(in0, in1, bb["w0"], bb["b0"], bb["w1"], bb["b1"], x)
        x = self._c(self._c(x, bb["d0"], ACT_GELU), bb["d2"])
        x2 = self._new(x, C)
        ops.groupnorm(x, 8, bb["gn_w"], bb["gn_b"], x2)
        del x
        f2 = self._c(self._c(x2, bb["c20"], ACT_GELU), bb["c22"], res=x2)
        del x2
        f4 = self._c(f2, bb["down"])
        self._tap("cnn_backbone.x4", f4); self._tap("cnn_backbone.x2", f2)
        f2_0 = f2[:B]

        # -- feature pyramid U-Net with relative-PE attention at 1/32 (unet.py:65-112)
        pe = ops.pe_tables(H // 32, W // 32, dev)
        py = self._unet(f4, P["pyr"], pe, "feat_pyramid")
        for i, t in enumerate(py):
            self._tap(f"feat_pyramid.{i}", t)

        # -- stacked multi-resolution transformer (stacked_MRT.py:156-166)
        z = py
        for mp in P["mrt"]:
            z = self._mrt(*z, mp)
        ftr = z[0]
        self._tap("transformer", ftr)

        # -- initial disparity by optimal transport over the all-pairs scan-line cost volume (submodules.py:203-243)
        cv = ops.empty((B, h4, w4, w4), f32, dev)
        if not ops.row_corr_tc(ftr, P["di_w"], P["di_b"], cv):       # fp16: LN + split-fp16 tcgen05 correlation (fp32-grade accuracy)
            fn = self._ln(ftr, P["di_w"], P["di_b"], out_dtype=f32)
            ops.row_corr(fn, cv)
            del fn
        disp = ops.empty((B, h4, w4), f32, dev)
        conf = ops.empty((B, h4, w4), f32, dev)
        occ = ops.empty((B, h4, w4), f32, dev)
        amax = ops.empty((B, h4, w4), torch.int32, dev)
        ops.sinkhorn_softargmax(cv, disp, conf, occ, amax, self.use_positivity, 3)
        if self.taps is not None:
            self.taps["disp_init.disp"] = disp.clone().unsqueeze(1); self.taps["disp_init.conf"] = conf.clone().unsqueeze(1)
            self.taps["disp_init.occ"] = occ.clone().unsqueeze(1); self.taps["disp_init.cv"] = cv.clone()
            self.taps["disp_init.argmax"] = amax.clone().long()

        ftr0, fpy0 = ftr[:B], py[0][:B]

        # -- global refiner (refinenet.py:59-73)
        gr = P["gr"]
        gaux = ops.zeros((B, h4, w4, 64), adt, dev)
        ops.global_refiner_prep(disp, conf, gaux[..., 0:2])
        feat = self._c(self._c(ftr0, gr["i0"], ACT_GELU, x2=gaux), gr["i2"])
        rf = self._unet(feat, gr["unet"])[0]
        upd = self._c(rf, gr["out"], out_dtype=f32)                    # already scaled by 1e2
        ops.global_refiner_merge(disp, conf, upd, self.use_positivity)  # in place on disp
        if self.taps is not None:
            self.taps["global_refiner"] = disp.clone().unsqueeze(1)

        # -- context / hidden state (s2m2.py:164-166)
        fus = self._fusion(ftr0, fpy0, P["ffl"])
        lr = P["lr"]
        xcat = ops.empty((B, h4, w4, 256 + C), adt, dev)                # [disp_feat 96 | corr1 64 | corr2 64 | ctx C | conf_occ 32]
        ctx0 = xcat[..., 224:224 + C]
        self._c(self._c(fus, P["ctx0"], ACT_GELU), P["ctx2"], out=ctx0)
        self._tap("ctx0", ctx0)
        hx = ops.empty((B, h4, w4, 2 * C), adt, dev)                    # GRU state | GRU input
        rh = ops.empty((B, h4, w4, C), adt, dev)                        # r * h
        hidden = hx[..., :C]
        ops.unary(ctx0, hidden, ACT_TANH)

        # -- iterative local refinement (refinenet.py:126-154, s2m2.py:175-180)
        small = ops.empty((B, h4, w4, 4), f32, dev)                     # disp/100 | conf_logit | occ_logit | 0
        corr = ops.empty((B, h4, w4, 18), f32, dev)                     # 9 taps level 1 | 9 taps level 2, both /16
        for it in range(self.refine_iter):
            ops.local_refiner_prep(disp, conf, occ, small)
            ops.cv_lookup(cv, disp, corr)
            self._c(self._c(small[..., 0:1], lr["df0"], ACT_GELU), lr["df2"], out=xcat[..., 0:96])
            self._c(self._c(corr[..., 0:9], lr["cf10"], ACT_GELU), lr["cf12"], out=xcat[..., 96:160])
            self._c(self._c(corr[..., 9:18], lr["cf20"], ACT_GELU), lr["cf22"], out=xcat[..., 160:224])
            self._c(self._c(small[..., 1:3], lr["co0"], ACT_GELU), lr["co2"], out=xcat[..., 224 + C:256 + C])
            x = self._c(self._c(xcat, lr["cat0"], ACT_GELU), lr["cat2"])
            xin = hx[..., C:]
            self._unet(x, lr["unet"], out0=xin)          # the U-Net's last ConvBlock writes straight into the GRU input slot [h | x]
            del x
            for zr, q in ((lr["zr1"], lr["q1"]), (lr["zr2"], lr["q2"])):     # ConvGRU (refinenet.py:22-36)
                if ops.conv_emodes(hx):
                    # [z | r*h] = sigmoid(conv([h, x])) with r * h formed in the epilogue (mode 1 on the r columns), then
                    # h = (1 - z) h + z tanh(conv([r*h, x])) as the q conv's epilogue (mode 3), written over h in place
                    zrt = self._new(hx, 2 * C)
                    if ops.conv2d(hx, zr, ACT_SIGMOID, zrt, res=hidden, emode=ops.EMODE_MUL, res_from=C):
                        if not ops.conv2d(zrt[..., C:], q, ACT_TANH, hidden, res=zrt[..., :C], x2=xin, res2=hidden, emode=ops.EMODE_GRU):
                            qt = self._c(zrt[..., C:], q, ACT_TANH, x2=xin)
                            ops.gru_update(zrt[..., :C], qt, hidden)
                        continue
                zrt = self._c(hx, zr, ACT_SIGMOID)
                ops.mul(zrt[..., C:], hidden, rh)
                qt = self._c(rh, q, ACT_TANH, x2=xin)                       # conv over cat([r*h, x]) reads the two sources in place
                ops.gru_update(zrt[..., :C], qt, hidden)                    # hidden = (1-z) h + z q, in place
            t = self._c(hidden, lr["upd0"], ACT_GELU)
            upd = self._c(t, lr["upd2"], out_dtype=f32)                    # [d disp | d conf logit | d occ logit | pad]
            ops.local_refiner_update(disp, conf, occ, small, upd, self.use_positivity)
            if self.taps is not None:
                self.taps[f"refiner.{it}.disp"] = disp.clone().unsqueeze(1)
                self.taps[f"refiner.{it}.conf"] = conf.clone().unsqueeze(1)
                self.taps[f"refiner.{it}.occ"] = occ.clone().unsqueeze(1)

        # -- convex 4x upsampling (submodules.py:110-115, s2m2.py:101-114,184-187)
        u4 = P["u4"]
        m4in = ops.empty((B, H // 2, W // 2, 128), adt, dev)
        self._c(hidden, u4["x"], out=m4in[..., :64])
        self._c(f2_0, u4["y"], out=m4in[..., 64:])
        m4 = self._c(self._c(m4in, u4["c0"], ACT_RELU), u4["c2"], out_dtype=f32)   # (B,H,W,16) logits, 9 used
        up = ops.empty((B, H, W, 4), f32, dev)                          # disp*4 | occ | conf | 0   (interleaved)
        ops.upsample4x_convex(disp, occ, conf, m4, up)
        if self.taps is not None:
            self.taps["up4x.disp"] = up[..., 0].clone().unsqueeze(1)
        del m4, m4in

        # -- edge-guided 1x filter (submodules.py:137-145, s2m2.py:116-133,190-193)
        u1 = P["u1"]
        m1in = ops.zeros((B, H, W, 64), adt, dev)                       # disp 16 | rgb 16 | ctx 16 | 16 zero pad
        ops.mask1x_inputs(up, img[:B], u1["wd"], u1["bd"], u1["wr"], u1["br"], m1in[..., 0:32])
        self._c(f2_0, u1["ctx"], out=m1in[..., 32:48])
        m1 = self._c(self._c(m1in, u1["c0"], ACT_RELU), u1["c2"], out_dtype=f32)
        sc = 2 if self.output_upsample else 1                            # s2m2.py:123-127,194-195
        if crop_hw is not None and not self.output_upsample:
            hc, wc = int(crop_hw[0]), int(crop_hw[1])
            if not (0 < hc <= H and 0 < wc <= W):
                raise ValueError(f"crop_hw {crop_hw} does not fit the {H}x{W} frame")
            out = ops.empty((3, B, 1, hc, wc), f32, dev)
            ops.filter1x_crop(up, m1, out, (hc, wc))
            return out[0], out[1], out[2]
        out = ops.empty((3, B, 1, sc * H, sc * W), f32, dev)
        ops.filter1x(up, m1, out, self.output_upsample)
        return out[0], out[1], out[2]
