"""Plain-torch restatement of every op of the ``s2m2_b200`` op interface.  TEST INFRASTRUCTURE ONLY.

Two uses:
  * on CPU: ``S2M2(...).ops = TorchOps()`` runs the product's exact kernel *schedule* (model.py)
    with these ops, which validates weight packing, buffer aliasing and op ordering against the oracle;
  * on the GPU: each CUDA kernel is compared one-to-one with the function of the same name here.

Tensors are NHWC views (last-dim stride 1, arbitrary pixel stride), like the CUDA ops take them.
"""
import math

import torch
import torch.nn.functional as F

from oracle.s2m2_oracle import pe_tables


def _act(x, act):
    if act == 1:
        return F.gelu(x)
    if act == 2:
        return F.relu(x)
    if act == 3:
        return torch.sigmoid(x)
    if act == 4:
        return torch.tanh(x)
    return x


class TorchOps:
    def __init__(self, compute_dtype=torch.float32):
        self.cd = compute_dtype
        self._pe = {}

    # allocation
    def empty(self, shape, dtype, device):
        return torch.full(shape, float("nan") if dtype.is_floating_point else -7, dtype=dtype, device=device)

    def zeros(self, shape, dtype, device):
        return torch.zeros(shape, dtype=dtype, device=device)

    def pe_tables(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._pe:
            px, py = pe_tables(h, w)
            self._pe[key] = (px.to(device).contiguous(), py.to(device).contiguous(), h, w)
        return self._pe[key]

    # GEMM-class
    EMODE_ADD, EMODE_MUL, EMODE_BLEND, EMODE_GRU = 0, 1, 2, 3
    fuse_epilogues = True            # CPU schedule tests exercise the fused gate / GRU epilogue route of the fp16 schedule

    def conv_emodes(self, x) -> bool:
        return self.fuse_epilogues and x.dtype == torch.float16

    def conv2d(self, x, cw, act, out, res=None, x2=None, res2=None, emode=0, res_from=0):
        xin = x if x2 is None else torch.cat([x, x2], dim=3)
        assert xin.shape[3] == cw.cin, (cw.name, xin.shape, cw.cin)
        w4 = cw.w.float().reshape(cw.cout, cw.kh, cw.kw, cw.cin).permute(0, 3, 1, 2)
        y = F.conv2d(xin.float().permute(0, 3, 1, 2), w4, cw.b, stride=cw.stride, padding=(cw.ph, cw.pw))
        y = _act(y, act).permute(0, 2, 3, 1)
        if cw.shuffle:
            n, h, w, c4 = y.shape
            c = c4 // 4
            y = y.reshape(n, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w, c)
        if emode == 0:
            if res is not None:
                y = y + res.float()
        elif emode == 1:           # y[..., n] = a * R1[..., n - res_from] for n >= res_from   (include/s2m2b200.h)
            y = torch.cat([y[..., :res_from], y[..., res_from:] * res.float()], dim=3)
        elif emode == 2:           # gate blend
            g = y.clamp(0.01, 0.99)
            y = g * res.float() + (1 - g) * res2.float()
        elif emode == 3:           # GRU state update
            z = res.float()
            y = z * y + (1 - z) * res2.float()
        else:
            raise ValueError(emode)
        assert out.shape == y.shape, (cw.name, out.shape, y.shape)
        out.copy_(y.to(out.dtype))
        return True

    def ffn(self, z, l0, l2):
        """The fused kernel's semantics (used by the GPU kernel test); returning False makes the schedule take the
        LayerNorm + two conv2d route on CPU."""
        return False

    def mlp2(self, x, res, out, l0, l2, act):
        """Returning False makes the schedule take the two conv2d launches on CPU."""
        return False

    def mlp2_reference(self, x, res, l0, l2, act):
        h = _act(F.linear(x.float(), l0.w.float(), l0.b), act).to(x.dtype).float()
        y = F.linear(h, l2.w.float(), l2.b)
        return (y + res.float() if res is not None else y).to(x.dtype)

    def ln_linear(self, z, lin, out):
        """Returning False makes the schedule take the LayerNorm + conv2d route on CPU."""
        return False

    def ln_linear_reference(self, z, lin):
        zn = F.layer_norm(z.float(), (z.shape[3],)).to(z.dtype).float()
        return F.linear(zn, lin.w.float(), lin.b).to(z.dtype)

    def ffn_reference(self, z, l0, l2):
        zn = F.layer_norm(z.float(), (z.shape[3],)).to(z.dtype).float()
        h = F.gelu(F.linear(zn, l0.w.float(), l0.b)).to(z.dtype).float()
        return (F.linear(h, l2.w.float(), l2.b) + z.float()).to(z.dtype)

    def attention(self, qkv, o, S, N, heads, kv_shift, pe=None, pesum=None):
        C = o.shape[3]
        d = C // heads
        t = qkv.float().reshape(S, N, 3, heads, d)
        q, k, v = t[:, :, 0].transpose(1, 2), t[:, :, 1].transpose(1, 2), t[:, :, 2].transpose(1, 2)  # (S,heads,N,d)
        if kv_shift:
            k, v = torch.roll(k, -kv_shift, 0), torch.roll(v, -kv_shift, 0)     # kv seq = (s + shift) % S
        attn = (torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5).softmax(-1)
        out = torch.matmul(attn, v)                                             # (S,heads,N,d)
        o.copy_(out.transpose(1, 2).reshape(o.shape).to(o.dtype))
        if pesum is not None:
            px, py, h, w = pe[:4]
            ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            xs, ys = xs.reshape(-1).to(px.device), ys.reshape(-1).to(px.device)
            full = 0.5 * torch.cat([px[xs[:, None] - xs[None, :] + w - 1], py[ys[:, None] - ys[None, :] + h - 1]], 2)
            ps = torch.einsum("bnij,ijc->bnic", attn, full)                     # (S,heads,N,32)
            pesum.copy_(ps.transpose(1, 2).to(pesum.dtype))

    def row_corr_tc(self, f, gamma, beta, cv):
        """Returning False makes the schedule take the layernorm + row_corr route on CPU."""
        return False

    def row_corr(self, f, cv):
        B = cv.shape[0]
        f0, f1 = f[:B].float(), f[B:].float()
        cv.copy_(torch.matmul(f0, f1.transpose(-1, -2)))

    # normalisation / resampling / elementwise
    def layernorm(self, x, out, gamma=None, beta=None):
        y = F.layer_norm(x.float(), (x.shape[3],), gamma, beta)
        out.copy_(y.to(out.dtype))

    def groupnorm(self, x, groups, gamma, beta, out):
        y = F.group_norm(x.float().permute(0, 3, 1, 2), groups, gamma, beta).permute(0, 2, 3, 1)
        out.copy_(y.to(out.dtype))

    def avgpool2(self, x, out):
        out.copy_(F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).to(out.dtype))

    def upsample2x(self, x, out):
        y = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False)
        out.copy_(y.permute(0, 2, 3, 1).to(out.dtype))

    def gate_blend(self, f, g, z0, z1, out):
        w = g.float().clamp(0.01, 0.99)
        out.copy_((f.float() + w * z0.float() + (1 - w) * z1.float()).to(out.dtype))

    def copy_channels(self, src, dst):
        dst.copy_(src.to(dst.dtype))

    def unary(self, x, out, act):
        out.copy_(_act(x.float(), act).to(out.dtype))

    def mul(self, a, b, out):
        out.copy_((a.float() * b.float()).to(out.dtype))

    def gru_update(self, z, q, h):
        zf = z.float()
        h.copy_(((1 - zf) * h.float() + zf * q.float()).to(h.dtype))

    def preprocess(self, left, right, out):
        x = torch.cat([left, right], 0)
        out.copy_(((x / 255.0 - 0.5) * 2).permute(0, 2, 3, 1).to(out.dtype))

    def stem(self, left, right, w0, b0, w1, b1, out):
        dev = left.device
        w0, b0, w1, b1 = (t.to(dev) for t in (w0, b0, w1, b1))       # weights arrive as host tensors
        x = ((torch.cat([left, right], 0).float() / 255.0 - 0.5) * 2)
        h = F.gelu(F.conv2d(x, w0[:, :, None, None], b0))
        y = F.conv2d(h, w1[:, :, None, None], b1).permute(0, 2, 3, 1)            # (2B,H,W,16)
        n, H, W, _ = y.shape
        if out.shape[1] == H // 2 and out.shape[3] == 64:                          # space-to-depth layout: [(py*2+px)*16 + c]
            y = y.reshape(n, H // 2, 2, W // 2, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(n, H // 2, W // 2, 64)
            out.copy_(y.to(out.dtype))
            return
        out.zero_()
        out[..., :16] = y.to(out.dtype)

    def mask1x_inputs(self, up, rgb, wd, bd, wr, br, out):
        wd, bd, wr, br = (t.to(up.device) for t in (wd, bd, wr, br))
        d = up[..., 0:1].float().permute(0, 3, 1, 2)
        a = F.relu(F.conv2d(d, wd.reshape(16, 3, 3, 1).permute(0, 3, 1, 2), bd, padding=1))
        r = F.relu(F.conv2d(rgb.float().permute(0, 3, 1, 2), wr.reshape(16, 3, 3, 3).permute(0, 3, 1, 2), br, padding=1))
        out[..., 0:16] = a.permute(0, 2, 3, 1).to(out.dtype)
        out[..., 16:32] = r.permute(0, 2, 3, 1).to(out.dtype)

    # matching
    def sinkhorn_softargmax(self, cv, disp, conf, occ, amax, use_positivity, iters):
        B, h, w, _ = cv.shape
        dev = cv.device
        mask = torch.triu(torch.ones(w, w, dtype=torch.bool, device=dev), diagonal=1) if use_positivity \
            else torch.zeros(w, w, dtype=torch.bool, device=dev)
        a = F.pad(cv.masked_fill(mask, -1e4), (0, 1, 0, 1))
        marg = torch.cat([torch.ones(w), torch.tensor([float(w)])]).to(dev) / (2 * w)
        lm = marg.log().reshape(1, 1, w + 1)

        def lse(x, dim):
            m = x.max(dim=dim, keepdim=True).values
            return (m + torch.log((x - m).exp().sum(dim=dim, keepdim=True).clamp(min=1e-30))).squeeze(dim)
        v = lm - lse(a, 2)
        u = lm - lse(a + v.unsqueeze(2), 3)
        for _ in range(iters - 1):
            v = lm - lse(a + u.unsqueeze(3), 2)
            u = lm - lse(a + v.unsqueeze(2), 3)
        prob = ((a + u.unsqueeze(3) + v.unsqueeze(2))[:, :, :-1, :-1] + math.log(2 * w)).exp().masked_fill(mask, 0)
        ind = prob.argmax(dim=3, keepdim=True)
        pp = F.pad(prob, (2, 2))
        cf = torch.zeros_like(prob[..., :1])
        cr = torch.zeros_like(cf)
        for k in range(5):
            wg = torch.gather(pp, 3, ind + k)
            cf = cf + wg
            cr = cr + wg * (ind + k - 2)
        cr = (cr + 1e-4) / (cf + 1e-4)
        xs = torch.arange(w, dtype=cv.dtype, device=dev).reshape(1, 1, w)
        disp.copy_(xs - cr.squeeze(3))
        conf.copy_(cf.squeeze(3))
        occ.copy_(prob.sum(3))
        amax.copy_(ind.squeeze(3).to(amax.dtype))

    def cv_lookup(self, cv, disp, out):
        B, h, w, w2 = cv.shape
        cv2 = F.avg_pool2d(cv.reshape(B * h * w, 1, 1, w2), kernel_size=[1, 2]).reshape(B, h, w, w2 // 2)
        dx = torch.arange(-4, 5, dtype=cv.dtype, device=cv.device)
        xs = torch.arange(w, dtype=cv.dtype, device=cv.device).reshape(1, 1, w, 1)
        d = disp.reshape(B, h, w, 1)

        def lerp(row, pos):
            n = row.shape[-1]
            p0 = torch.floor(pos)
            t = pos - p0
            i0 = p0.long()
            i1 = i0 + 1
            v0 = torch.gather(row, 3, i0.clamp(0, n - 1)) * ((i0 >= 0) & (i0 <= n - 1))
            v1 = torch.gather(row, 3, i1.clamp(0, n - 1)) * ((i1 >= 0) & (i1 <= n - 1))
            return (1 - t) * v0 + t * v1
        out[..., 0:9] = (lerp(cv, xs - d + dx) / 16).to(out.dtype)
        out[..., 9:18] = (lerp(cv2, xs / 2 - d / 2 + dx) / 16).to(out.dtype)

    # refiner glue
    def global_refiner_prep(self, disp, conf, out):
        mask = (conf > 0.2).float()
        out[..., 0] = (disp / 1e2 * mask).to(out.dtype)
        out[..., 1] = (mask * conf).logit(eps=1e-1).to(out.dtype)

    def global_refiner_merge(self, disp, conf, upd, use_positivity):
        mask = (conf > 0.2).float()
        d = mask * disp + (1 - mask) * upd[..., 0].float()
        if use_positivity:
            d = d.clamp(min=0)
        disp.copy_(d)

    def local_refiner_prep(self, disp, conf, occ, small):
        small[..., 0] = disp / 1e2
        small[..., 1] = conf.logit(eps=1e-2)
        small[..., 2] = occ.logit(eps=1e-2)
        small[..., 3] = 0

    def local_refiner_update(self, disp, conf, occ, small, upd, use_positivity):
        d = disp + upd[..., 0].float()
        if use_positivity:
            d = d.clamp(min=0)
        w = disp.shape[-1]
        xs = torch.arange(w, dtype=torch.float32, device=disp.device).reshape(1, 1, w)
        conf.copy_(torch.sigmoid(upd[..., 1].float() + small[..., 1]))
        occ.copy_(torch.sigmoid(upd[..., 2].float() + small[..., 2]) * (xs - d >= 0))
        disp.copy_(d)

    # upsampling
    @staticmethod
    def _unfold(x):   # (B,1,H,W) -> (B,9,H,W) replicate pad, row-major (dy,dx)
        B, C, H, W = x.shape
        xp = F.pad(x, (1, 1, 1, 1), mode="replicate")
        return torch.cat([xp[:, :, i:i + H, j:j + W] for i in range(3) for j in range(3)], dim=1)

    def upsample4x_convex(self, disp, occ, conf, logits, up):
        B, h, w = disp.shape
        wts = logits[..., :9].float().permute(0, 3, 1, 2).softmax(dim=1)
        for c, m in enumerate((disp * 4, occ, conf)):
            xu = F.interpolate(self._unfold(m.unsqueeze(1)), (4 * h, 4 * w), mode="nearest")
            up[..., c] = (xu * wts).sum(1)
        up[..., 3] = 0

    def filter1x(self, up, logits, out, upsample2=False):
        lg = logits[..., :9].float().permute(0, 3, 1, 2)
        if upsample2:
            lg = F.interpolate(lg, scale_factor=2, mode="bilinear", align_corners=False)
        wts = lg.softmax(dim=1)
        for c in range(3):
            xu = self._unfold(up[..., c].unsqueeze(1))
            if upsample2:
                xu = F.interpolate(xu, scale_factor=2, mode="nearest")
            out[c] = (xu * wts).sum(1, keepdim=True) * (2.0 if (upsample2 and c == 0) else 1.0)
