// Host side of the forward in C++ behind the C-ABI: a model handle that takes the reference's state-dict
// entries, packs them into the layouts the kernels read, and runs S2M2.forward (s2m2.py:136-197) as a schedule of this
// library's ops on a caller-provided stream and workspace.
//
//   s2m2b200_create -> s2m2b200_set_param (x every state-dict key) -> s2m2b200_finalize_params
//   s2m2b200_workspace_bytes(B, H, W, ...)   walks the schedule once in plan mode (common.cuh): every op runs its checks and
//                                            dispatch, nothing is enqueued; the arena's high-water mark is the answer
//   s2m2b200_forward(...)                    the same walk with real launches; no allocation, no synchronisation, capturable
//
// The schedule is the one s2m2_b200/model.py runs through ctypes (the nn.Module surface the reference's Python callers see);
// the two are held together by tests: the op traces are compared line by line on the CPU (tests/test_host_forward_cpu.py) and
// the outputs bit for bit on the GPU (tests/test_host_forward_gpu.py).  Comments cite the reference lines each block follows.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace s2m2 {
namespace host {

// ================================================================================================
// small helpers
// ================================================================================================
struct Fail : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] static void fail(const char* fmt, ...) {
  char b[600];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof(b), fmt, ap);
  va_end(ap);
  throw Fail(b);
}
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static bool env_on(const char* name) {   // "1" unless set to "0" (the switches of ops.py)
  const char* e = getenv(name);
  return !(e && e[0] == '0' && e[1] == 0);
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// fp32 host tensor (a state-dict entry or an intermediate of the packing), up to 4-D, row-major
struct HT {
  std::vector<float> v;
  int64_t s[4] = {1, 1, 1, 1};
  int nd = 0;
  HT() {}
  HT(int64_t a, int64_t b, int64_t c, int64_t d) : v((size_t)(a * b * c * d), 0.f), nd(4) { s[0] = a; s[1] = b; s[2] = c; s[3] = d; }
  float& at(int64_t a, int64_t b, int64_t c, int64_t d) { return v[(size_t)(((a * s[1] + b) * s[2] + c) * s[3] + d)]; }
  float at(int64_t a, int64_t b, int64_t c, int64_t d) const { return v[(size_t)(((a * s[1] + b) * s[2] + c) * s[3] + d)]; }
};

// ================================================================================================
// packed weights
// ================================================================================================
struct ConvW {   // one implicit-GEMM layer: w is (cout, kh*kw*cin), K order tap-major / channel-minor (model.py ConvW)
  std::string name;
  size_t w_off = 0, b_off = 0;
  bool has_b = false;
  int kh = 1, kw = 1, stride = 1, ph = 0, pw = 0, cin = 0, cout = 0, shuffle = 0;
  const void* w = nullptr;     // resolved when the blob has its final address
  const float* b = nullptr;
};
struct HostConv {   // a layer before it is committed to the blob (so that several can be concatenated along cout)
  std::vector<float> w, b;
  bool has_b = false;
  int kh = 1, kw = 1, stride = 1, ph = 0, pw = 0, cin = 0, cout = 0, shuffle = 0;
};
struct Attn { ConvW qkv, proj, pe_proj; bool has_pe = false; int heads = 1; };
struct Ffn { ConvW l0, l2; };
struct Basic { Attn cross, self; Ffn ffn_c, ffn; };
struct Global { Attn self, cross; Ffn ffn, ffn_c; bool has_cross = false; };
struct ConvBlock { ConvW c0, c2, x0, x2; };
struct Fusion { ConvW gf0, g2, f2; int dim = 0; };
struct UpDown { ConvW down[3], up[3]; };
struct Unet { UpDown ud; Fusion cat[3]; ConvBlock enc[3], dec[3]; std::vector<Global> enc3s, dec3s; };
struct Mrt { UpDown ud; Fusion dcat[4], ucat[3]; Basic enc[3], dec[3]; std::vector<Global> enc3s, dec3s; };

struct Entry { std::string name; size_t off, bytes; int dtype; };   // blob directory (dtype: S2M2_F32 / S2M2_F16)

struct Knobs {   // the switches and thresholds of ops.CudaOps, same environment variables
  int conv_impl, ffn_min_tokens, ffn_min_tokens_wide, mlp2_min_tokens, mlp2_max_tokens, ln_linear_min_tokens;
  bool ffn_fused, row_corr_fused, fuse_epilogues, mlp2_fused, ln_linear_fused, attn_tc;
  Knobs() {
    conv_impl = env_int("S2M2_CONV_IMPL", 0);
    ffn_fused = env_on("S2M2_FFN_FUSED");
    ffn_min_tokens = env_int("S2M2_FFN_MIN_TOKENS", 16384);
    ffn_min_tokens_wide = env_int("S2M2_FFN_MIN_TOKENS_WIDE", 37888);
    row_corr_fused = env_on("S2M2_ROWCORR_TC");
    fuse_epilogues = env_on("S2M2_FUSE_EPILOGUES");
    mlp2_fused = env_on("S2M2_MLP2_FUSED");
    mlp2_min_tokens = env_int("S2M2_MLP2_MIN_TOKENS", 16384);
    mlp2_max_tokens = env_int("S2M2_MLP2_MAX_TOKENS", 120000);
    ln_linear_fused = env_on("S2M2_LNL_FUSED");
    ln_linear_min_tokens = env_int("S2M2_LNL_MIN_TOKENS", 16384);
    attn_tc = env_on("S2M2_ATTN_TC");
  }
};

struct ShapeKey {
  int B, H, W, u8, ch, cw;
  bool operator<(const ShapeKey& o) const {
    return memcmp(this, &o, sizeof(*this)) < 0;
  }
};
struct Plan {   // what s2m2b200_workspace_bytes leaves behind for one (B, H, W, input type, crop)
  size_t workspace = 0;
  // constants of the shape: the relative-position tables of the 1/32-resolution grid and the indicator matrix of the tcgen05 PE
  // attention (attentions.py:42-48,165-226), plus the token buffer of the scan-line attention's ABI.  Device memory owned by the
  // model (host memory in a device-less plan).
  char* consts = nullptr;
  bool consts_on_device = false;
  size_t pe_x = 0, pe_y = 0, pe_ind = 0, attn_work = 0, consts_bytes = 0, ind_bytes = 0;
  bool has_ind = false;
  // graph mode (s2m2b200_set_graph_mode): the forward of this shape captured once for one set of buffer addresses
  cudaGraphExec_t exec = nullptr;
  const void *g_left = nullptr, *g_right = nullptr, *g_out = nullptr, *g_work = nullptr;
};

}  // namespace host
}  // namespace s2m2

using namespace s2m2;
using namespace s2m2::host;

struct S2m2Model {
  int C = 0, NTR = 0, use_positivity = 0, output_upsample = 0, refine_iter = 3, prec = S2M2_F16;
  std::map<std::string, HT> params;
  bool finalized = false;
  Knobs knobs;
  // packed blob
  std::vector<Entry> dir;
  std::vector<uint8_t> staging;     // host image of the blob while packing
  char* blob = nullptr;             // final address: device memory, or 256-aligned host memory on a box without a device
  bool blob_on_device = false;
  size_t blob_bytes = 0;
  // packed tree (model.py _pack)
  std::vector<float> bb_w0, bb_b0, bb_w1, bb_b1, u1_wd, u1_bd, u1_wr, u1_br;   // tiny weights that travel as kernel parameters
  ConvW bb_d0, bb_d2, bb_c20, bb_c22, bb_down;
  size_t gn_w = 0, gn_b = 0, di_w = 0, di_b = 0;
  Unet pyr;
  std::vector<Mrt> mrt;
  ConvW gr_i0, gr_i2, gr_out;
  Unet gr_unet;
  Fusion ffl;
  ConvW ctx0, ctx2;
  ConvW lr_df0, lr_df2, lr_cf10, lr_cf12, lr_cf20, lr_cf22, lr_co0, lr_co2, lr_cat0, lr_cat2, lr_upd0, lr_upd2, lr_zr1, lr_q1, lr_zr2, lr_q2;
  Unet lr_unet;
  ConvW u4_x, u4_y, u4_c0, u4_c2, u1_ctx, u1_c0, u1_c2;
  std::vector<ConvW*> all_convs;
  std::map<ShapeKey, Plan> plans;
  bool graph_mode = false;
  cudaStream_t cap_stream = nullptr;   // graph mode captures on a stream of its own (the caller's may be the legacy default stream)

  ~S2m2Model() {
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (blob) {
      if (blob_on_device) cudaFree(blob);
      else free(blob);
    }
    for (auto& kv : plans) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (!kv.second.consts) continue;
      if (kv.second.consts_on_device) cudaFree(kv.second.consts);
      else free(kv.second.consts);
    }
  }
};

namespace s2m2 {
namespace host {

// ================================================================================================
// packing (model.py:119-334)
// ================================================================================================
struct Packer {
  S2m2Model& m;
  explicit Packer(S2m2Model& mm) : m(mm) {}

  bool has(const std::string& n) const { return m.params.count(n) != 0; }
  const HT& get(const std::string& n) const {
    auto it = m.params.find(n);
    if (it == m.params.end()) fail("finalize_params: parameter '%s' was never set", n.c_str());
    return it->second;
  }
  const HT& get4(const std::string& n, int64_t a, int64_t b, int64_t c, int64_t d) const {   // a dimension < 0 is free
    const HT& t = get(n);
    const int64_t want[4] = {a, b, c, d};
    bool ok = t.nd == 4;
    for (int i = 0; ok && i < 4; ++i) ok = want[i] < 0 || t.s[i] == want[i];
    if (!ok) fail("finalize_params: '%s' has shape (%lld,%lld,%lld,%lld)/%dd, expected (%lld,%lld,%lld,%lld)", n.c_str(), (long long)t.s[0],
                  (long long)t.s[1], (long long)t.s[2], (long long)t.s[3], t.nd, (long long)a, (long long)b, (long long)c, (long long)d);
    return t;
  }
  std::vector<float> vec(const std::string& n, int64_t len) const {
    const HT& t = get(n);
    if (t.nd != 1 || (len >= 0 && t.s[0] != len)) fail("finalize_params: '%s' must be a vector of %lld", n.c_str(), (long long)len);
    return t.v;
  }
  HT weight4(const std::string& n) const {   // conv weights as they are, linear weights (cout, cin) as 1x1 convolutions
    const HT& t = get(n);
    if (t.nd == 4) return t;
    if (t.nd != 2) fail("finalize_params: '%s' must be 2-D or 4-D", n.c_str());
    HT r = t;
    r.nd = 4; r.s[0] = t.s[0]; r.s[1] = t.s[1]; r.s[2] = 1; r.s[3] = 1;
    return r;
  }

  // ---- blob ----
  size_t add_raw(const std::string& name, const void* data, size_t bytes, int dtype) {
    const size_t off = align_up(m.staging.size(), 256);
    m.staging.resize(off + bytes);
    memcpy(m.staging.data() + off, data, bytes);
    m.dir.push_back({name, off, bytes, dtype});
    return off;
  }
  size_t add_f32(const std::string& name, const std::vector<float>& v) { return add_raw(name, v.data(), v.size() * 4, S2M2_F32); }
  size_t add_act(const std::string& name, const std::vector<float>& v) {   // in the activation dtype of the precision mode
    if (m.prec == S2M2_F32) return add_f32(name, v);
    std::vector<__half> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(v[i]);   // round-to-nearest-even, as torch's .to(float16)
    return add_raw(name, h.data(), h.size() * 2, S2M2_F16);
  }

  // ---- model.py _mk ----
  HostConv mk(HT w4, const std::vector<float>* b, int stride = 1, int ph = -1, int pw = -1, bool shuffle = false, int cin_to = 0, int cout_to = 0) const {
    std::vector<float> bias;
    if (b) bias = *b;
    if (cin_to > w4.s[1]) {
      HT p(w4.s[0], cin_to, w4.s[2], w4.s[3]);
      for (int64_t o = 0; o < w4.s[0]; ++o)
        for (int64_t c = 0; c < w4.s[1]; ++c)
          for (int64_t y = 0; y < w4.s[2]; ++y)
            for (int64_t x = 0; x < w4.s[3]; ++x) p.at(o, c, y, x) = w4.at(o, c, y, x);
      w4 = std::move(p);
    }
    if (cout_to > w4.s[0]) {
      HT p(cout_to, w4.s[1], w4.s[2], w4.s[3]);
      memcpy(p.v.data(), w4.v.data(), w4.v.size() * 4);
      w4 = std::move(p);
      if (b) bias.resize((size_t)cout_to, 0.f);
    }
    HostConv c;
    c.cout = (int)w4.s[0]; c.cin = (int)w4.s[1]; c.kh = (int)w4.s[2]; c.kw = (int)w4.s[3];
    c.stride = stride;
    c.ph = ph < 0 ? c.kh / 2 : ph;
    c.pw = pw < 0 ? c.kw / 2 : pw;
    c.shuffle = shuffle ? 1 : 0;
    c.w.resize(w4.v.size());
    const int64_t K = (int64_t)c.kh * c.kw * c.cin;
    for (int64_t o = 0; o < c.cout; ++o)
      for (int64_t y = 0; y < c.kh; ++y)
        for (int64_t x = 0; x < c.kw; ++x)
          for (int64_t ci = 0; ci < c.cin; ++ci) c.w[(size_t)(o * K + (y * c.kw + x) * c.cin + ci)] = w4.at(o, ci, y, x);
    c.has_b = b != nullptr;
    c.b = std::move(bias);
    return c;
  }
  HostConv conv_h(const std::string& name, int stride = 1, int ph = -1, int pw = -1, int cin_to = 0, int cout_to = 0) const {
    HT w = weight4(name + ".weight");
    if (has(name + ".bias")) {
      std::vector<float> b = vec(name + ".bias", w.s[0]);
      return mk(std::move(w), &b, stride, ph, pw, false, cin_to, cout_to);
    }
    return mk(std::move(w), nullptr, stride, ph, pw, false, cin_to, cout_to);
  }
  void commit(ConvW& dst, const HostConv& c, const std::string& name) {
    dst.name = name;
    dst.kh = c.kh; dst.kw = c.kw; dst.stride = c.stride; dst.ph = c.ph; dst.pw = c.pw; dst.cin = c.cin; dst.cout = c.cout; dst.shuffle = c.shuffle;
    dst.w_off = add_act("w:" + name, c.w);
    dst.has_b = c.has_b;
    if (c.has_b) dst.b_off = add_f32("b:" + name, c.b);
    m.all_convs.push_back(&dst);
  }
  void conv(ConvW& dst, const std::string& name, int stride = 1, int ph = -1, int pw = -1, int cin_to = 0, int cout_to = 0) {
    commit(dst, conv_h(name, stride, ph, pw, cin_to, cout_to), name);
  }
  // model.py _cat_out: layers sharing one input as a single GEMM (rows concatenated; missing biases are zeros)
  void cat_out(ConvW& dst, const std::vector<HostConv>& parts, const std::string& name) {
    HostConv c = parts[0];
    c.w.clear(); c.b.clear(); c.cout = 0;
    for (const HostConv& p : parts) {
      c.w.insert(c.w.end(), p.w.begin(), p.w.end());
      if (p.has_b) c.b.insert(c.b.end(), p.b.begin(), p.b.end());
      else c.b.insert(c.b.end(), (size_t)p.cout, 0.f);
      c.cout += p.cout;
    }
    c.has_b = true;
    c.shuffle = 0;
    commit(dst, c, name);
  }
  // model.py _convT2: ConvTranspose2d(k=2,s=2), weight (Cin,Cout,2,2) -> GEMM rows (a*2+b)*Cout + co, pixel-shuffle store
  void convT2(ConvW& dst, const std::string& name, int cout_to = 0) {
    const HT& w = get4(name + ".weight", -1, -1, 2, 2);
    const int64_t cin = w.s[0], co_real = w.s[1];
    const int64_t cout = cout_to > co_real ? cout_to : co_real;
    std::vector<float> b = vec(name + ".bias", co_real);
    b.resize((size_t)cout, 0.f);
    HT g(4 * cout, cin, 1, 1);
    for (int64_t a = 0; a < 2; ++a)
      for (int64_t bb = 0; bb < 2; ++bb)
        for (int64_t co = 0; co < co_real; ++co)
          for (int64_t ci = 0; ci < cin; ++ci) g.at((a * 2 + bb) * cout + co, ci, 0, 0) = w.at(ci, co, a, bb);
    std::vector<float> b4;
    for (int r = 0; r < 4; ++r) b4.insert(b4.end(), b.begin(), b.end());
    commit(dst, mk(std::move(g), &b4, 1, -1, -1, true), name);
  }
  // model.py _convT1: ConvTranspose2d(k=1) = 1x1 convolution with the weight's first two dims swapped
  void convT1(ConvW& dst, const std::string& name, int cin_to, int cout_to) {
    const HT& w = get4(name + ".weight", -1, -1, 1, 1);
    HT t(w.s[1], w.s[0], 1, 1);
    for (int64_t i = 0; i < w.s[0]; ++i)
      for (int64_t o = 0; o < w.s[1]; ++o) t.at(o, i, 0, 0) = w.at(i, o, 0, 0);
    std::vector<float> b = vec(name + ".bias", w.s[1]);
    commit(dst, mk(std::move(t), &b, 1, -1, -1, false, cin_to, cout_to), name);
  }
  // model.py _pack_d0: conv1_down.0 (16 -> 64, 5x5, stride 2; submodules.py:74); in fp16 the stem writes space-to-depth pixels and
  // the layer becomes a dense 3x3 stride-1 convolution over 64 channels
  void pack_d0(ConvW& dst, const std::string& name) {
    const HT& w = get4(name + ".weight", -1, 16, 5, 5);
    std::vector<float> b = vec(name + ".bias", w.s[0]);
    if (m.prec != S2M2_F16) {
      commit(dst, mk(w, &b, 2), name);
      return;
    }
    HT w3(w.s[0], 64, 3, 3);
    for (int ty = 0; ty < 3; ++ty)
      for (int py = 0; py < 2; ++py)
        for (int tx = 0; tx < 3; ++tx)
          for (int px = 0; px < 2; ++px) {
            const int ky = 2 * ty + py, kx = 2 * tx + px;
            if (ky >= 5 || kx >= 5) continue;
            const int c0 = (py * 2 + px) * 16;
            for (int64_t o = 0; o < w.s[0]; ++o)
              for (int c = 0; c < 16; ++c) w3.at(o, c0 + c, ty, tx) = w.at(o, c, ky, kx);
          }
    commit(dst, mk(std::move(w3), &b, 1), name);
  }
  int head_pad() const { return m.prec == S2M2_F16 ? 16 : 0; }   // model.py _head_pad

  void attn(Attn& a, const std::string& p, int heads, bool pe) {
    cat_out(a.qkv, {conv_h(p + ".q"), conv_h(p + ".k"), conv_h(p + ".v")}, p + ".qkv");
    conv(a.proj, p + ".proj");
    a.heads = heads;
    a.has_pe = pe;
    if (pe) conv(a.pe_proj, p + ".pe_proj");
  }
  void ffn(Ffn& f, const std::string& p) {
    conv(f.l0, p + ".ffn.0");
    conv(f.l2, p + ".ffn.2");
  }
  void basic(Basic& b, const std::string& p, int heads) {
    attn(b.cross, p + ".cross_attn.attn", heads, false);
    ffn(b.ffn_c, p + ".ffn_c");
    attn(b.self, p + ".self_attn.attn", heads, false);
    ffn(b.ffn, p + ".ffn");
  }
  void global(Global& g, const std::string& p, int heads, bool cross, bool pe) {
    attn(g.self, p + ".self_attn.attn", heads, pe);
    ffn(g.ffn, p + ".ffn");
    g.has_cross = cross;
    if (cross) {
      attn(g.cross, p + ".cross_attn.attn", heads, false);
      ffn(g.ffn_c, p + ".ffn_c");
    }
  }
  void convblock(ConvBlock& c, const std::string& p) {
    conv(c.c0, p + ".convs.0");
    conv(c.c2, p + ".convs.2");
    conv(c.x0, p + ".convs_1x.0");
    conv(c.x2, p + ".convs_1x.2");
  }
  void fusion(Fusion& f, const std::string& p) {
    HostConv g0 = conv_h(p + ".feature_gate.0"), f0 = conv_h(p + ".feature_fusion.0");
    f.dim = g0.cout;
    cat_out(f.gf0, {g0, f0}, p + ".gf0");
    conv(f.g2, p + ".feature_gate.2");
    conv(f.f2, p + ".feature_fusion.2");
  }
  void updown(UpDown& u, const std::string& p) {
    for (int i = 0; i < 3; ++i) conv(u.down[i], p + ".down_conv" + std::to_string(i) + ".1");
    for (int i = 0; i < 3; ++i) conv(u.up[i], p + ".up_conv" + std::to_string(i) + ".1");
  }
  void unet(Unet& u, const std::string& p, int n_attn, bool pe) {
    updown(u.ud, p);
    for (int i = 0; i < 3; ++i) {
      const std::string si = std::to_string(i);
      fusion(u.cat[i], p + ".concat_conv" + si);
      convblock(u.enc[i], p + ".enc" + si);
      convblock(u.dec[i], p + ".dec" + si);
    }
    u.enc3s.resize((size_t)n_attn);
    u.dec3s.resize((size_t)n_attn);
    for (int i = 0; i < n_attn; ++i) global(u.enc3s[(size_t)i], p + ".enc3s." + std::to_string(i), 8, false, pe);
    for (int i = 0; i < n_attn; ++i) global(u.dec3s[(size_t)i], p + ".dec3s." + std::to_string(i), 8, false, false);
  }
  void mrt(Mrt& t, const std::string& p) {
    updown(t.ud, p);
    for (int i = 1; i <= 3; ++i) fusion(t.dcat[i], p + ".down_concat" + std::to_string(i));
    for (int i = 0; i < 3; ++i) {
      const std::string si = std::to_string(i);
      fusion(t.ucat[i], p + ".up_concat" + si);
      basic(t.enc[i], p + ".enc_attn" + si, 1 << i);
      basic(t.dec[i], p + ".dec_attn" + si, 1 << i);
    }
    t.enc3s.resize(2);
    t.dec3s.resize(2);
    for (int i = 0; i < 2; ++i) global(t.enc3s[(size_t)i], p + ".enc_attn3s." + std::to_string(i), 8, true, false);
    for (int i = 0; i < 2; ++i) global(t.dec3s[(size_t)i], p + ".dec_attn3s." + std::to_string(i), 8, true, false);
  }

  void pack_all() {
    const int C = m.C;
    const std::string b = "cnn_backbone";
    {   // stem (submodules.py:70-72): two 1x1 layers, weights as kernel parameters
      const HT& w0 = get4(b + ".conv0.0.weight", 16, 3, 1, 1);
      const HT& w1 = get4(b + ".conv0.2.weight", 16, 16, 1, 1);
      m.bb_w0 = w0.v; m.bb_b0 = vec(b + ".conv0.0.bias", 16);
      m.bb_w1 = w1.v; m.bb_b1 = vec(b + ".conv0.2.bias", 16);
    }
    pack_d0(m.bb_d0, b + ".conv1_down.0");
    conv(m.bb_d2, b + ".conv1_down.2");
    m.gn_w = add_f32("bb.gn_w", vec(b + ".norm1.weight", C));
    m.gn_b = add_f32("bb.gn_b", vec(b + ".norm1.bias", C));
    conv(m.bb_c20, b + ".conv2.0");
    conv(m.bb_c22, b + ".conv2.2");
    conv(m.bb_down, b + ".conv2_down.0", 2);
    unet(m.pyr, "feat_pyramid", 2 * m.NTR, true);
    m.mrt.resize((size_t)m.NTR);
    for (int l = 0; l < m.NTR; ++l) mrt(m.mrt[(size_t)l], "transformer.uformer_list." + std::to_string(l));
    m.di_w = add_f32("di_w", vec("disp_init.layer_norm.weight", C));
    m.di_b = add_f32("di_b", vec("disp_init.layer_norm.bias", C));
    {   // global refiner (refinenet.py:43-73): input order [disp, conf_logit, ctx(C)] -> two sources [ctx(C) | aux(64)]; the *1e2 of
        // line 71 folded into the head
      const std::string g = "global_refiner";
      const HT& w = get4(g + ".init_feat.0.weight", -1, C + 2, 3, 3);
      HT wp(w.s[0], C + 64, 3, 3);
      for (int64_t o = 0; o < w.s[0]; ++o)
        for (int y = 0; y < 3; ++y)
          for (int x = 0; x < 3; ++x) {
            for (int c = 0; c < C; ++c) wp.at(o, c, y, x) = w.at(o, 2 + c, y, x);
            wp.at(o, C, y, x) = w.at(o, 0, y, x);
            wp.at(o, C + 1, y, x) = w.at(o, 1, y, x);
          }
      std::vector<float> bi = vec(g + ".init_feat.0.bias", w.s[0]);
      commit(m.gr_i0, mk(std::move(wp), &bi), g + ".init_feat.0");
      conv(m.gr_i2, g + ".init_feat.2");
      unet(m.gr_unet, g + ".refine_unet", 1, false);
      HT ow = weight4(g + ".out_feat.0.weight");
      for (float& v : ow.v) v = v * 100.0f;
      std::vector<float> ob = vec(g + ".out_feat.0.bias", ow.s[0]);
      for (float& v : ob) v = v * 100.0f;
      commit(m.gr_out, mk(std::move(ow), &ob, 1, -1, -1, false, 0, head_pad()), g + ".out_feat.0");
    }
    fusion(m.ffl, "feat_fusion_layer");
    conv(m.ctx0, "ctx_feat.0");
    conv(m.ctx2, "ctx_feat.2");
    {
      const std::string r = "refiner";
      conv(m.lr_df0, r + ".disp_feat.0", 1, -1, -1, 0, 128);
      conv(m.lr_df2, r + ".disp_feat.2", 1, -1, -1, 128, 0);
      conv(m.lr_cf10, r + ".corr_feat1.0", 1, -1, -1, 0, 128);
      conv(m.lr_cf12, r + ".corr_feat1.2", 1, -1, -1, 128, 0);
      conv(m.lr_cf20, r + ".corr_feat2.0", 1, -1, -1, 0, 128);
      conv(m.lr_cf22, r + ".corr_feat2.2", 1, -1, -1, 128, 0);
      conv(m.lr_co0, r + ".conf_occ_feat.0");
      conv(m.lr_co2, r + ".conf_occ_feat.2");
      conv(m.lr_cat0, r + ".disp_corr_ctx_cat.0");
      conv(m.lr_cat2, r + ".disp_corr_ctx_cat.2");
      unet(m.lr_unet, r + ".refine_unet", 1, false);
      cat_out(m.lr_upd0, {conv_h(r + ".disp_update.0"), conv_h(r + ".conf_occ_update.0")}, r + ".upd0");
      {   // model.py _pack_update_heads: disp_update.2 (C->1) and conf_occ_update.2 (C->2) as one block-structured 3-channel head
        const HT& wd = get4(r + ".disp_update.2.weight", 1, C, 3, 3);
        const HT& wc = get4(r + ".conf_occ_update.2.weight", 2, C, 3, 3);
        HT w(3, 2 * C, 3, 3);
        for (int c = 0; c < C; ++c)
          for (int y = 0; y < 3; ++y)
            for (int x = 0; x < 3; ++x) {
              w.at(0, c, y, x) = wd.at(0, c, y, x);
              w.at(1, C + c, y, x) = wc.at(0, c, y, x);
              w.at(2, C + c, y, x) = wc.at(1, c, y, x);
            }
        commit(m.lr_upd2, mk(std::move(w), nullptr, 1, -1, -1, false, 0, head_pad()), r + ".update_heads");
      }
      cat_out(m.lr_zr1, {conv_h(r + ".gru.convz1", 1, 1, 0), conv_h(r + ".gru.convr1", 1, 1, 0)}, "zr1");
      conv(m.lr_q1, r + ".gru.convq1", 1, 1, 0);
      cat_out(m.lr_zr2, {conv_h(r + ".gru.convz2", 1, 0, 1), conv_h(r + ".gru.convr2", 1, 0, 1)}, "zr2");
      conv(m.lr_q2, r + ".gru.convq2", 1, 0, 1);
    }
    {
      const std::string u = "upsample_mask_4x_refine";
      convT2(m.u4_x, u + ".conv_x");
      conv(m.u4_y, u + ".conv_y");
      conv(m.u4_c0, u + ".conv_concat.0");
      convT2(m.u4_c2, u + ".conv_concat.2", 16);
    }
    {
      const std::string u = "upsample_mask_1x";
      // ConvTranspose2d(k=3,p=1) = 3x3 convolution with flipped taps; [co][tap] and [co][tap*3+ci] (submodules.py:127-131)
      const HT& wd = get4(u + ".conv_disp.0.weight", 1, 16, 3, 3);
      const HT& wr = get4(u + ".conv_rgb.0.weight", 3, 16, 3, 3);
      m.u1_wd.assign(16 * 9, 0.f);
      m.u1_wr.assign(16 * 27, 0.f);
      for (int co = 0; co < 16; ++co)
        for (int y = 0; y < 3; ++y)
          for (int x = 0; x < 3; ++x) {
            m.u1_wd[(size_t)(co * 9 + y * 3 + x)] = wd.at(0, co, 2 - y, 2 - x);
            for (int ci = 0; ci < 3; ++ci) m.u1_wr[(size_t)(co * 27 + (y * 3 + x) * 3 + ci)] = wr.at(ci, co, 2 - y, 2 - x);
          }
      m.u1_bd = vec(u + ".conv_disp.0.bias", 16);
      m.u1_br = vec(u + ".conv_rgb.0.bias", 16);
      convT2(m.u1_ctx, u + ".conv_ctx");
      conv(m.u1_c0, u + ".conv_concat.0", 1, -1, -1, 64, 64);
      convT1(m.u1_c2, u + ".conv_concat.2", 64, 16);
    }
  }
};

// ================================================================================================
// workspace arena and tensor views
// ================================================================================================
struct Arena {
  char* base = nullptr;
  size_t cap = 0, top = 0, peak = 0;
  std::map<size_t, size_t> free_;   // offset -> size, coalesced
  size_t alloc(size_t bytes) {
    bytes = align_up(bytes ? bytes : 1, 256);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second < bytes) continue;
      const size_t off = it->first, rest = it->second - bytes;
      free_.erase(it);
      if (rest) free_[off + bytes] = rest;
      return off;
    }
    const size_t off = top;
    top += bytes;
    if (top > peak) peak = top;
    if (cap && top > cap) fail("forward: workspace too small (%zu bytes needed so far, %zu given); size it with s2m2b200_workspace_bytes", top, cap);
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = align_up(bytes ? bytes : 1, 256);
    auto it = free_.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_.erase(nx);
    }
    if (it != free_.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_.erase(it);
        it = pv;
      }
    }
    if (it->first + it->second == top) {   // the last block is free again: give it back to the bump region
      top = it->first;
      free_.erase(it);
    }
  }
};
struct Blk {
  Arena* a;
  size_t off, bytes;
  Blk(Arena* aa, size_t o, size_t b) : a(aa), off(o), bytes(b) {}
  ~Blk() {
    if (trace_on()) {
      char t[64];
      snprintf(t, sizeof(t), "free %p", (void*)(a->base + off));
      trace_append(t);
    }
    a->release(off, bytes);
  }
};
struct T {   // strided-NHWC view (ops.tview): channels contiguous, pixels ld elements apart
  char* p = nullptr;
  int dt = S2M2_F16, n = 0, h = 0, w = 0, c = 0;
  int64_t ld = 0;
  std::shared_ptr<Blk> blk;
  bool ok() const { return p != nullptr; }
  size_t es() const { return dt == S2M2_F16 ? 2 : 4; }
  int64_t pixels() const { return (int64_t)n * h * w; }
  S2m2Tensor view() const {
    S2m2Tensor t;
    t.ptr = p; t.dtype = dt; t.N = n; t.H = h; t.W = w; t.C = c; t.ld = ld;
    return t;
  }
  T ch(int c0, int c1) const {   // x[..., c0:c1]
    T r = *this;
    r.p = p + (size_t)c0 * es();
    r.c = c1 - c0;
    return r;
  }
  T first(int n1) const {   // x[:n1]
    T r = *this;
    r.n = n1;
    return r;
  }
  T as(int n2, int h2, int w2, int c2) const {   // reshape of a contiguous tensor
    if (ld != c || (int64_t)n2 * h2 * w2 * c2 != pixels() * c) fail("forward: reshape of a non-contiguous view");
    T r = *this;
    r.n = n2; r.h = h2; r.w = w2; r.c = c2; r.ld = c2;
    return r;
  }
};
static S2m2Tensor null_view() {
  S2m2Tensor t;
  memset(&t, 0, sizeof(t));
  return t;
}

// ================================================================================================
// the schedule (model.py:344-695)
// ================================================================================================
struct Run {
  S2m2Model& m;
  const Knobs& k;
  Arena arena;
  void* stream;
  bool plan;
  const Plan* pl = nullptr;
  int adt;

  Run(S2m2Model& mm, void* s, bool is_plan) : m(mm), k(mm.knobs), stream(s), plan(is_plan), adt(mm.prec) {}

  void chk(int rc, const char* what) {
    if (rc != 0) fail("%s failed (%d): %s", what, rc, s2m2b200_last_error());
  }
  void note_alloc(const void* p, size_t bytes) {
    if (!trace_on()) return;
    char b[96];
    snprintf(b, sizeof(b), "alloc %p %zu", p, bytes);
    trace_append(b);
  }
  T empty(int n, int h, int w, int c, int dt) {
    T t;
    t.dt = dt; t.n = n; t.h = h; t.w = w; t.c = c; t.ld = c;
    const size_t bytes = (size_t)n * h * w * c * t.es();
    const size_t off = arena.alloc(bytes);
    t.blk = std::make_shared<Blk>(&arena, off, bytes);
    t.p = arena.base + off;
    note_alloc(t.p, bytes);
    return t;
  }
  T zeros(int n, int h, int w, int c, int dt) {
    T t = empty(n, h, w, c, dt);
    if (!plan) {
      cudaError_t e = cudaMemsetAsync(t.p, 0, (size_t)n * h * w * c * t.es(), (cudaStream_t)stream);
      if (e != cudaSuccess) fail("forward: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    }
    return t;
  }
  T like(const T& ref, int c, int dt = -1, int h = -1, int w = -1) {   // model.py _new
    return empty(ref.n, h < 0 ? ref.h : h, w < 0 ? ref.w : w, c, dt < 0 ? adt : dt);
  }

  // ---- ops.CudaOps, one method each ----
  bool conv2d(const T& x, const ConvW& cw, int act, const T& out, const T* res = nullptr, const T* x2 = nullptr, const T* res2 = nullptr, int emode = 0,
              int res_from = 0) {
    S2m2Conv a;
    memset(&a, 0, sizeof(a));
    a.x = x.view();
    a.x2 = x2 ? x2->view() : null_view();
    a.y = out.view();
    a.res = res ? res->view() : null_view();
    a.res2 = res2 ? res2->view() : null_view();
    a.emode = emode;
    a.res_from = res_from;
    const int cin = x.c + (x2 ? x2->c : 0);
    if (cin != cw.cin) fail("conv %s: input has %d channels, weights expect %d", cw.name.c_str(), cin, cw.cin);
    a.w = cw.w;
    a.w_dtype = adt;
    a.bias = cw.has_b ? cw.b : nullptr;
    a.KH = cw.kh; a.KW = cw.kw; a.stride = cw.stride; a.pad_h = cw.ph; a.pad_w = cw.pw;
    a.Cout = cw.cout; a.act = act; a.shuffle2x2 = cw.shuffle;
    a.batch = 1;
    a.impl = k.conv_impl;
    const int rc = s2m2b200_conv2d(&a, stream);
    if (rc == 1 && emode != 0) return false;
    if (rc != 0) fail("conv2d[%s] failed (%d): %s", cw.name.c_str(), rc, s2m2b200_last_error());
    return true;
  }
  // model.py _c
  T conv(const T& x, const ConvW& cw, int act = S2M2_ACT_NONE, const T* out = nullptr, const T* res = nullptr, const T* x2 = nullptr, int out_dtype = -1) {
    T o;
    if (out) o = *out;
    else if (cw.shuffle) o = like(x, cw.cout / 4, out_dtype, 2 * x.h, 2 * x.w);
    else o = like(x, cw.cout, out_dtype, (x.h + 2 * cw.ph - cw.kh) / cw.stride + 1, (x.w + 2 * cw.pw - cw.kw) / cw.stride + 1);
    conv2d(x, cw, act, o, res, x2);
    return o;
  }
  T ln(const T& x, const float* gamma = nullptr, const float* beta = nullptr, int out_dtype = -1) {
    T o = like(x, x.c, out_dtype);
    S2m2Tensor xv = x.view(), ov = o.view();
    chk(s2m2b200_layernorm(&xv, &ov, gamma, beta, stream), "layernorm");
    return o;
  }
  bool ffn_fused(const T& z, const ConvW& l0, const ConvW& l2) {   // ops.ffn
    if (!k.ffn_fused || !l0.has_b || !l2.has_b || l0.cin != l0.cout || l2.cin != l2.cout) return false;
    if (z.pixels() < (z.c <= 256 ? k.ffn_min_tokens : k.ffn_min_tokens_wide)) return false;
    S2m2Tensor zv = z.view();
    const int rc = s2m2b200_ffn(&zv, &zv, l0.w, l0.b, l2.w, l2.b, adt, stream);
    if (rc == 1) return false;
    chk(rc, "ffn");
    return true;
  }
  bool mlp2(const T& x, const T& res, const T& out, const ConvW& l0, const ConvW& l2, int act) {   // ops.mlp2
    if (!k.mlp2_fused || !l0.has_b || !l2.has_b || l0.cin != l0.cout || l2.cin != l2.cout || l0.kh != 1 || l2.kh != 1) return false;
    const int64_t ntok = x.pixels();
    if (ntok < k.mlp2_min_tokens || ntok > k.mlp2_max_tokens) return false;
    S2m2Tensor xv = x.view(), rv = res.view(), ov = out.view();
    const int rc = s2m2b200_mlp2(&xv, &rv, &ov, l0.w, l0.b, l2.w, l2.b, adt, 0, act, stream);
    if (rc == 1) return false;
    chk(rc, "mlp2");
    return true;
  }
  bool ln_linear(const T& z, const ConvW& lin, const T& out) {   // ops.ln_linear
    if (!k.ln_linear_fused || lin.kh != 1 || lin.kw != 1 || lin.shuffle) return false;
    if (z.pixels() < k.ln_linear_min_tokens) return false;
    S2m2Tensor zv = z.view(), ov = out.view();
    const int rc = s2m2b200_ln_linear(&zv, &ov, lin.w, lin.has_b ? lin.b : nullptr, adt, stream);
    if (rc == 1) return false;
    chk(rc, "ln_linear");
    return true;
  }
  bool conv_emodes(const T& x) const { return k.fuse_epilogues && x.dt == S2M2_F16 && k.conv_impl != 1; }
  void attention(const T& qkv, const T& o, int S, int N, int heads, int kv_shift, bool pe, const T* pesum) {   // ops.attention
    S2m2Tensor q = qkv.view(), ov = o.view();
    if (pesum) {
      S2m2Tensor ps;
      ps.ptr = pesum->p; ps.dtype = pesum->dt; ps.N = 1; ps.H = 1; ps.W = S * N * heads; ps.C = 32; ps.ld = 32;
      const bool ind = pl->has_ind && k.attn_tc;
      chk(s2m2b200_attention(&q, &ov, S, N, heads, kv_shift, (const float*)(pl->consts + pl->pe_x), (const float*)(pl->consts + pl->pe_y), pe_h, pe_w, &ps,
                             ind ? pl->consts + pl->pe_ind : nullptr, ind ? (int64_t)pl->ind_bytes : 0, stream),
          "attention");
      return;
    }
    (void)pe;
    void* work = nullptr;
    int64_t wbytes = 0;
    if (k.attn_tc) {
      wbytes = s2m2b200_attention_workspace_bytes(S, N, heads, o.c / heads, qkv.dt);
      if (wbytes > 0) work = pl->consts + pl->attn_work;
    }
    chk(s2m2b200_attention(&q, &ov, S, N, heads, kv_shift, nullptr, nullptr, 0, 0, nullptr, work, wbytes, stream), "attention");
  }
  int pe_h = 0, pe_w = 0;

  // ---- model.py building blocks ----
  T ffn(const T& z, const Ffn& p) {   // attentions.py:245-250
    if (ffn_fused(z, p.l0, p.l2)) return z;
    T h = conv(ln(z), p.l0, S2M2_ACT_GELU);
    conv(h, p.l2, S2M2_ACT_NONE, &z, &z);
    return z;
  }
  // LN -> fused QKV GEMM -> attention -> proj + residual (attentions.py:33-54,77-96,119-161)
  T attn(const T& z, const Attn& p, int S, int N, bool cross, const T* out, bool pe) {
    const int C = z.c, heads = p.heads;
    T qkv = like(z, p.qkv.cout);
    if (!ln_linear(z, p.qkv, qkv)) conv(ln(z), p.qkv, S2M2_ACT_NONE, &qkv);
    T o = like(z, C);
    T pesum;
    const bool with_pe = pe && p.has_pe;
    if (with_pe) pesum = empty(S, N, heads, 32, adt);
    attention(qkv, o, S, N, heads, cross ? S / 2 : 0, with_pe, with_pe ? &pesum : nullptr);
    if (with_pe) {   // out_head += pe_proj(pe_sum) (attentions.py:47-48)
      const int d = C / heads;
      T ov = o.as(1, 1, S * N * heads, d);
      T pv = pesum.as(1, 1, S * N * heads, 32);
      conv(pv, p.pe_proj, S2M2_ACT_NONE, &ov, &ov);
    }
    T dst = out ? *out : z;
    conv(o, p.proj, S2M2_ACT_NONE, &dst, &z);
    return dst;
  }
  T basic_block(T z, const Basic& p, bool fresh) {   // attentions.py:347-355
    T out = fresh ? like(z, z.c) : z;
    const int n = z.n, h = z.h, w = z.w;
    z = attn(z, p.cross, n * h, w, true, &out, false);
    z = ffn(z, p.ffn_c);
    z = attn(z, p.self, n * h, w, false, nullptr, false);
    return ffn(z, p.ffn);
  }
  T global_block(T z, const Global& p, bool fresh, bool pe) {   // attentions.py:311-321
    T out = fresh ? like(z, z.c) : z;
    const int n = z.n, hw = z.h * z.w;
    if (p.has_cross) {
      z = attn(z, p.cross, n, hw, true, &out, false);
      z = ffn(z, p.ffn_c);
      z = attn(z, p.self, n, hw, false, nullptr, pe);
    } else {
      z = attn(z, p.self, n, hw, false, &out, pe);
    }
    return ffn(z, p.ffn);
  }
  T convblock(const T& z, const ConvBlock& p, const T* out = nullptr) {   // attentions.py:277-281
    T a = conv(conv(z, p.c0, S2M2_ACT_GELU), p.c2);
    T o = out ? *out : a;
    if (mlp2(z, a, o, p.x0, p.x2, S2M2_ACT_RELU)) return o;
    T u = conv(z, p.x0, S2M2_ACT_RELU);
    return conv(u, p.x2, S2M2_ACT_NONE, out, &a);
  }
  T fusion(const T& z0, const T& z1, const Fusion& p) {   // feature_fusion.py:24-33
    const int d = p.dim;
    T t = conv(z0, p.gf0, S2M2_ACT_GELU, nullptr, nullptr, &z1);
    if (conv_emodes(t)) {
      T bl = like(z0, d);
      if (conv2d(t.ch(0, d), p.g2, S2M2_ACT_SIGMOID, bl, &z0, nullptr, &z1, 2)) return conv(t.ch(d, t.c), p.f2, S2M2_ACT_NONE, nullptr, &bl);
    }
    T g = conv(t.ch(0, d), p.g2, S2M2_ACT_SIGMOID);
    T f = conv(t.ch(d, t.c), p.f2);
    S2m2Tensor fv = f.view(), gv = g.view(), a = z0.view(), b = z1.view();
    chk(s2m2b200_gate_blend(&fv, &gv, &a, &b, &fv, stream), "gate_blend");
    return f;
  }
  T down(const T& z, const ConvW& cw) {   // unet.py:25-30
    T p = like(z, z.c, -1, z.h / 2, z.w / 2);
    S2m2Tensor zv = z.view(), pv = p.view();
    chk(s2m2b200_avgpool2(&zv, &pv, stream), "avgpool2");
    return conv(p, cw);
  }
  T up(const T& z, const ConvW& cw) {   // unet.py:32-37, as 1x1 conv -> bilinear x2
    T t = conv(z, cw);
    T u = like(t, t.c, -1, 2 * t.h, 2 * t.w);
    S2m2Tensor tv = t.view(), uv = u.view();
    chk(s2m2b200_upsample2x(&tv, &uv, stream), "upsample2x");
    return u;
  }
  struct Pyr { T z0, z1, z2, z3; };
  Pyr unet(const T& z, const Unet& p, bool pe, const T* out0 = nullptr) {   // unet.py:65-112
    T z0 = convblock(z, p.enc[0]);
    T z1 = convblock(down(z0, p.ud.down[0]), p.enc[1]);
    T z2 = convblock(down(z1, p.ud.down[1]), p.enc[2]);
    T z3 = down(z2, p.ud.down[2]);
    for (const Global& g : p.enc3s) z3 = global_block(z3, g, false, pe);
    for (const Global& g : p.dec3s) z3 = global_block(z3, g, false, false);
    T z2n = convblock(fusion(z2, up(z3, p.ud.up[2]), p.cat[2]), p.dec[2]);
    T z1n = convblock(fusion(z1, up(z2n, p.ud.up[1]), p.cat[1]), p.dec[1]);
    T z0n = convblock(fusion(z0, up(z1n, p.ud.up[0]), p.cat[0]), p.dec[0], out0);
    return {z0n, z1n, z2n, z3};
  }
  Pyr mrt(Pyr z, const Mrt& p) {   // stacked_MRT.py:89-133
    T z0 = basic_block(z.z0, p.enc[0], true);
    T z1 = basic_block(fusion(z.z1, down(z0, p.ud.down[0]), p.dcat[1]), p.enc[1], false);
    T z2 = basic_block(fusion(z.z2, down(z1, p.ud.down[1]), p.dcat[2]), p.enc[2], false);
    T z3 = fusion(z.z3, down(z2, p.ud.down[2]), p.dcat[3]);
    for (const Global& g : p.enc3s) z3 = global_block(z3, g, false, false);
    for (const Global& g : p.dec3s) z3 = global_block(z3, g, false, false);
    z2 = basic_block(fusion(z2, up(z3, p.ud.up[2]), p.ucat[2]), p.dec[2], false);
    z1 = basic_block(fusion(z1, up(z2, p.ud.up[1]), p.ucat[1]), p.dec[1], false);
    z0 = basic_block(fusion(z0, up(z1, p.ud.up[0]), p.ucat[0]), p.dec[0], false);
    return {z0, z1, z2, z3};
  }

  const float* f32p(size_t off) const { return (const float*)(m.blob + off); }

  // model.py _forward_eager (s2m2.py:136-197)
  void forward(const void* in0, const void* in1, bool u8, int B, int H, int W, int crop_h, int crop_w, float* out) {
    const int C = m.C, f32 = S2M2_F32;
    const int h4 = H / 4, w4 = W / 4;
    const bool pos = m.use_positivity != 0;
    // -- normalise + NCHW -> NHWC, left and right stacked on the batch dim (s2m2.py:86-96)
    T img = empty(2 * B, H, W, 3, adt);
    {
      S2m2Tensor iv = img.view();
      if (u8) chk(s2m2b200_preprocess_u8((const uint8_t*)in0, (const uint8_t*)in1, B, H, W, &iv, stream), "preprocess");
      else chk(s2m2b200_preprocess((const float*)in0, (const float*)in1, B, H, W, &iv, stream), "preprocess");
    }
    // -- CNN backbone (submodules.py:88-93)
    T f2, f4;
    {
      T x = adt == S2M2_F16 ? empty(2 * B, H / 2, W / 2, 64, adt) : empty(2 * B, H, W, 16, adt);
      S2m2Tensor xv = x.view();
      if (u8) chk(s2m2b200_stem_u8((const uint8_t*)in0, (const uint8_t*)in1, B, H, W, m.bb_w0.data(), m.bb_b0.data(), m.bb_w1.data(), m.bb_b1.data(), &xv, stream), "stem");
      else chk(s2m2b200_stem((const float*)in0, (const float*)in1, B, H, W, m.bb_w0.data(), m.bb_b0.data(), m.bb_w1.data(), m.bb_b1.data(), &xv, stream), "stem");
      x = conv(conv(x, m.bb_d0, S2M2_ACT_GELU), m.bb_d2);
      T x2 = like(x, C);
      {   // ops.groupnorm: scratch for the per-block partial sums
        const int nblk = (x.h * x.w + 255) / 256;
        const size_t wbytes = (size_t)2 * x.n * 8 * (1 + nblk) * sizeof(double);
        const size_t off = arena.alloc(wbytes);
        Blk work(&arena, off, wbytes);
        note_alloc(arena.base + off, wbytes);
        S2m2Tensor a = x.view(), b = x2.view();
        chk(s2m2b200_groupnorm(&a, &b, 8, f32p(m.gn_w), f32p(m.gn_b), (double*)(arena.base + off), stream), "groupnorm");
      }
      x = T();
      f2 = conv(conv(x2, m.bb_c20, S2M2_ACT_GELU), m.bb_c22, S2M2_ACT_NONE, nullptr, &x2);
      x2 = T();
      f4 = conv(f2, m.bb_down);
    }
    T f2_0 = f2.first(B);

    // -- feature pyramid U-Net with relative-PE attention at 1/32 (unet.py:65-112)
    pe_h = H / 32; pe_w = W / 32;
    Pyr py = unet(f4, m.pyr, true);
    // -- stacked multi-resolution transformer (stacked_MRT.py:156-166)
    Pyr z = py;
    for (const Mrt& mp : m.mrt) z = mrt(z, mp);
    T ftr = z.z0;

    // -- initial disparity by optimal transport over the all-pairs scan-line cost volume (submodules.py:203-243)
    T cv = empty(B, h4, w4, w4, f32);
    bool rc_done = false;
    if (k.row_corr_fused && ftr.dt == S2M2_F16) {   // ops.row_corr_tc
      const int64_t nb = s2m2b200_row_corr_workspace_bytes(B, h4, w4, ftr.c);
      if (nb > 0) {
        const size_t off = arena.alloc((size_t)nb);
        Blk work(&arena, off, (size_t)nb);
        note_alloc(arena.base + off, (size_t)nb);
        S2m2Tensor fv = ftr.view();
        const int rc = s2m2b200_row_corr(&fv, f32p(m.di_w), f32p(m.di_b), (float*)cv.p, arena.base + off, nb, stream);
        if (rc != 1) {
          chk(rc, "row_corr");
          rc_done = true;
        }
      }
    }
    if (!rc_done) {   // LayerNorm -> one fp32 GEMM per scan-line (ops.row_corr)
      T fn = ln(ftr, f32p(m.di_w), f32p(m.di_b), f32);
      S2m2Conv a;
      memset(&a, 0, sizeof(a));
      a.x.ptr = fn.p; a.x.dtype = f32; a.x.N = 1; a.x.H = 1; a.x.W = w4; a.x.C = fn.c; a.x.ld = fn.c;
      a.y.ptr = cv.p; a.y.dtype = f32; a.y.N = 1; a.y.H = 1; a.y.W = w4; a.y.C = w4; a.y.ld = w4;
      a.w = fn.p + (size_t)B * h4 * w4 * fn.c * 4;
      a.w_dtype = f32;
      a.KH = a.KW = a.stride = 1;
      a.Cout = w4;
      a.batch = B * h4;
      a.x_bs = (int64_t)w4 * fn.c; a.w_bs = (int64_t)w4 * fn.c; a.y_bs = (int64_t)w4 * w4;
      a.impl = 1;
      chk(s2m2b200_conv2d(&a, stream), "row_corr");
    }
    T disp = empty(B, h4, w4, 1, f32), conf = empty(B, h4, w4, 1, f32), occ = empty(B, h4, w4, 1, f32);
    {
      T amax = empty(B, h4, w4, 1, f32);   // int32 indices
      chk(s2m2b200_sinkhorn_softargmax((const float*)cv.p, B, h4, w4, (float*)disp.p, (float*)conf.p, (float*)occ.p, (int32_t*)amax.p, pos, 3, stream),
          "sinkhorn_softargmax");
    }
    float *dp = (float*)disp.p, *cp = (float*)conf.p, *op = (float*)occ.p;
    const int64_t npx = (int64_t)B * h4 * w4;
    T ftr0 = ftr.first(B), fpy0 = py.z0.first(B);

    // -- global refiner (refinenet.py:59-73)
    {
      T gaux = zeros(B, h4, w4, 64, adt);
      S2m2Tensor gv = gaux.ch(0, 2).view();
      chk(s2m2b200_global_refiner_prep(dp, cp, npx, &gv, stream), "global_refiner_prep");
      T feat = conv(conv(ftr0, m.gr_i0, S2M2_ACT_GELU, nullptr, nullptr, &gaux), m.gr_i2);
      T rf = unet(feat, m.gr_unet, false).z0;
      T upd = conv(rf, m.gr_out, S2M2_ACT_NONE, nullptr, nullptr, nullptr, f32);   // already scaled by 1e2
      chk(s2m2b200_global_refiner_merge(dp, cp, (const float*)upd.p, upd.c, npx, pos, stream), "global_refiner_merge");
    }

    // -- context / hidden state (s2m2.py:164-166)
    T fus = fusion(ftr0, fpy0, m.ffl);
    T xcat = empty(B, h4, w4, 256 + C, adt);   // [disp_feat 96 | corr1 64 | corr2 64 | ctx C | conf_occ 32]
    T ctx0 = xcat.ch(224, 224 + C);
    conv(conv(fus, m.ctx0, S2M2_ACT_GELU), m.ctx2, S2M2_ACT_NONE, &ctx0);
    T hx = empty(B, h4, w4, 2 * C, adt);       // GRU state | GRU input
    T rh = empty(B, h4, w4, C, adt);           // r * h
    T hidden = hx.ch(0, C);
    {
      S2m2Tensor a = ctx0.view(), b = hidden.view();
      chk(s2m2b200_unary(&a, &b, S2M2_ACT_TANH, stream), "unary");
    }

    // -- iterative local refinement (refinenet.py:126-154, s2m2.py:175-180)
    T small = empty(B, h4, w4, 4, f32);        // disp/100 | conf_logit | occ_logit | 0
    T corr = empty(B, h4, w4, 18, f32);        // 9 taps level 1 | 9 taps level 2, both /16
    for (int it = 0; it < m.refine_iter; ++it) {
      chk(s2m2b200_local_refiner_prep(dp, cp, op, npx, (float*)small.p, stream), "local_refiner_prep");
      chk(s2m2b200_cv_lookup((const float*)cv.p, dp, B, h4, w4, (float*)corr.p, stream), "cv_lookup");
      {
        T o0 = xcat.ch(0, 96), o1 = xcat.ch(96, 160), o2 = xcat.ch(160, 224), o3 = xcat.ch(224 + C, 256 + C);
        conv(conv(small.ch(0, 1), m.lr_df0, S2M2_ACT_GELU), m.lr_df2, S2M2_ACT_NONE, &o0);
        conv(conv(corr.ch(0, 9), m.lr_cf10, S2M2_ACT_GELU), m.lr_cf12, S2M2_ACT_NONE, &o1);
        conv(conv(corr.ch(9, 18), m.lr_cf20, S2M2_ACT_GELU), m.lr_cf22, S2M2_ACT_NONE, &o2);
        conv(conv(small.ch(1, 3), m.lr_co0, S2M2_ACT_GELU), m.lr_co2, S2M2_ACT_NONE, &o3);
      }
      T xin = hx.ch(C, 2 * C);
      {
        T x = conv(conv(xcat, m.lr_cat0, S2M2_ACT_GELU), m.lr_cat2);
        unet(x, m.lr_unet, false, &xin);   // the U-Net's last ConvBlock writes straight into the GRU input slot [h | x]
      }
      const ConvW* zrs[2] = {&m.lr_zr1, &m.lr_zr2};
      const ConvW* qs[2] = {&m.lr_q1, &m.lr_q2};
      for (int half = 0; half < 2; ++half) {   // ConvGRU (refinenet.py:22-36)
        const ConvW &zr = *zrs[half], &q = *qs[half];
        if (conv_emodes(hx)) {
          T zrt = like(hx, 2 * C);
          if (conv2d(hx, zr, S2M2_ACT_SIGMOID, zrt, &hidden, nullptr, nullptr, 1, C)) {
            T zt = zrt.ch(0, C), rt = zrt.ch(C, 2 * C);
            if (!conv2d(rt, q, S2M2_ACT_TANH, hidden, &zt, &xin, &hidden, 3)) {
              T qt = conv(rt, q, S2M2_ACT_TANH, nullptr, nullptr, &xin);
              S2m2Tensor a = zt.view(), b = qt.view(), c = hidden.view();
              chk(s2m2b200_gru_update(&a, &b, &c, stream), "gru_update");
            }
            continue;
          }
        }
        T zrt = conv(hx, zr, S2M2_ACT_SIGMOID);
        {
          S2m2Tensor a = zrt.ch(C, 2 * C).view(), b = hidden.view(), c = rh.view();
          chk(s2m2b200_mul(&a, &b, &c, stream), "mul");
        }
        T qt = conv(rh, q, S2M2_ACT_TANH, nullptr, nullptr, &xin);
        S2m2Tensor a = zrt.ch(0, C).view(), b = qt.view(), c = hidden.view();
        chk(s2m2b200_gru_update(&a, &b, &c, stream), "gru_update");
      }
      T t = conv(hidden, m.lr_upd0, S2M2_ACT_GELU);
      T upd = conv(t, m.lr_upd2, S2M2_ACT_NONE, nullptr, nullptr, nullptr, f32);   // [d disp | d conf logit | d occ logit | pad]
      if (upd.c < 3) fail("forward: update head has %d channels", upd.c);
      chk(s2m2b200_local_refiner_update(dp, cp, op, (const float*)small.p, (const float*)upd.p, upd.c, B, h4, w4, pos, stream), "local_refiner_update");
    }

    // -- convex 4x upsampling (submodules.py:110-115, s2m2.py:101-114,184-187)
    T up4 = T();
    {
      T m4in = empty(B, H / 2, W / 2, 128, adt);
      T a = m4in.ch(0, 64), b = m4in.ch(64, 128);
      conv(hidden, m.u4_x, S2M2_ACT_NONE, &a);
      conv(f2_0, m.u4_y, S2M2_ACT_NONE, &b);
      T m4 = conv(conv(m4in, m.u4_c0, S2M2_ACT_RELU), m.u4_c2, S2M2_ACT_NONE, nullptr, nullptr, nullptr, f32);   // (B,H,W,16) logits, 9 used
      up4 = empty(B, H, W, 4, f32);   // disp*4 | occ | conf | 0
      if (m4.c < 9) fail("forward: 4x mask head has %d channels", m4.c);
      chk(s2m2b200_upsample4x_convex(dp, op, cp, (const float*)m4.p, m4.c, B, h4, w4, (float*)up4.p, stream), "upsample4x_convex");
    }

    // -- edge-guided 1x filter (submodules.py:137-145, s2m2.py:116-133,190-193)
    T m1in = zeros(B, H, W, 64, adt);   // disp 16 | rgb 16 | ctx 16 | 16 zero pad
    {
      S2m2Tensor rgb = img.first(B).view(), o = m1in.ch(0, 32).view();
      chk(s2m2b200_mask1x_inputs((const float*)up4.p, &rgb, m.u1_wd.data(), m.u1_bd.data(), m.u1_wr.data(), m.u1_br.data(), &o, stream), "mask1x_inputs");
    }
    {
      T c = m1in.ch(32, 48);
      conv(f2_0, m.u1_ctx, S2M2_ACT_NONE, &c);
    }
    T m1 = conv(conv(m1in, m.u1_c0, S2M2_ACT_RELU), m.u1_c2, S2M2_ACT_NONE, nullptr, nullptr, nullptr, f32);
    if (m1.c < 9) fail("forward: 1x mask head has %d channels", m1.c);
    const bool crop = crop_h > 0 && crop_w > 0 && !m.output_upsample;
    const int sc = m.output_upsample ? 2 : 1;
    note_alloc(out, (size_t)3 * B * (crop ? (size_t)crop_h * crop_w : (size_t)sc * H * sc * W) * 4);
    if (crop) {
      if (crop_h > H || crop_w > W) fail("forward: crop %dx%d does not fit the %dx%d frame", crop_h, crop_w, H, W);
      chk(s2m2b200_filter1x_crop((const float*)up4.p, (const float*)m1.p, m1.c, B, H, W, crop_h, crop_w, out, stream), "filter1x_crop");
    } else {
      chk(s2m2b200_filter1x((const float*)up4.p, (const float*)m1.p, m1.c, B, H, W, out, m.output_upsample, stream), "filter1x");
    }
  }
};

// a pointer that is never dereferenced, for walking the schedule without memory behind it
static char* const kFakeBase = (char*)(uintptr_t)0x100000000000ull;
static char* const kFakeIO = (char*)(uintptr_t)0x200000000000ull;

static bool have_device() {
  int n = 0;
  const bool ok = cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
  (void)cudaGetLastError();
  return ok;
}

static void resolve(S2m2Model& m) {
  for (ConvW* c : m.all_convs) {
    c->w = m.blob + c->w_off;
    c->b = c->has_b ? (const float*)(m.blob + c->b_off) : nullptr;
  }
}

static Plan& make_plan(S2m2Model& m, const ShapeKey& key) {
  auto it = m.plans.find(key);
  if (it != m.plans.end()) return it->second;
  const int B = key.B, H = key.H, W = key.W;
  Plan pl;
  // constants of the shape
  const int h = H / 32, w = W / 32;
  std::vector<float> px((size_t)(2 * w + 1) * 16), py((size_t)(2 * h + 1) * 16);
  if (s2m2b200_pe_table(w, px.data()) != 0 || s2m2b200_pe_table(h, py.data()) != 0) fail("workspace_bytes: pe_table failed: %s", s2m2b200_last_error());
  std::vector<uint8_t> ind;
  if (h + w <= 80) {   // ops.pe_tables
    ind.resize((size_t)s2m2b200_pe_indicator_bytes(h, w));
    if (s2m2b200_pe_indicator(h, w, ind.data()) != 0) fail("workspace_bytes: pe_indicator failed: %s", s2m2b200_last_error());
    pl.has_ind = true;
    pl.ind_bytes = ind.size();
  }
  pl.pe_x = 0;
  pl.pe_y = align_up(pl.pe_x + px.size() * 4, 256);
  pl.pe_ind = align_up(pl.pe_y + py.size() * 4, 256);
  pl.attn_work = align_up(pl.pe_ind + ind.size(), 256);
  pl.consts_bytes = pl.attn_work + 256;
  std::vector<uint8_t> hostc(pl.consts_bytes, 0);
  memcpy(hostc.data() + pl.pe_x, px.data(), px.size() * 4);
  memcpy(hostc.data() + pl.pe_y, py.data(), py.size() * 4);
  if (!ind.empty()) memcpy(hostc.data() + pl.pe_ind, ind.data(), ind.size());
  if (m.blob_on_device) {
    if (cudaMalloc((void**)&pl.consts, pl.consts_bytes) != cudaSuccess) fail("workspace_bytes: cudaMalloc(%zu) failed", pl.consts_bytes);
    pl.consts_on_device = true;
    if (cudaMemcpy(pl.consts, hostc.data(), pl.consts_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(pl.consts);
      fail("workspace_bytes: upload of the shape constants failed");
    }
  } else {
    pl.consts = (char*)aligned_alloc(256, align_up(pl.consts_bytes, 256));
    memcpy(pl.consts, hostc.data(), pl.consts_bytes);
  }
  if (trace_on()) {
    char b[160];
    snprintf(b, sizeof(b), "static pe_x %p %zu", (void*)(pl.consts + pl.pe_x), px.size() * 4); trace_append(b);
    snprintf(b, sizeof(b), "static pe_y %p %zu", (void*)(pl.consts + pl.pe_y), py.size() * 4); trace_append(b);
    if (pl.has_ind) { snprintf(b, sizeof(b), "static pe_ind %p %zu", (void*)(pl.consts + pl.pe_ind), ind.size()); trace_append(b); }
    snprintf(b, sizeof(b), "static attn_work %p %d", (void*)(pl.consts + pl.attn_work), 256); trace_append(b);
  }
  // walk the schedule with nothing behind the pointers
  Run run(m, nullptr, true);
  run.pl = &pl;
  run.arena.base = kFakeBase;
  const size_t in_bytes = (size_t)B * 3 * H * W * (key.u8 ? 1 : 4);
  const int prev = s2m2b200_set_plan_mode(1);
  try {
    if (trace_on()) {
      char b[160];
      snprintf(b, sizeof(b), "static in0 %p %zu", (void*)kFakeIO, in_bytes); trace_append(b);
      snprintf(b, sizeof(b), "static in1 %p %zu", (void*)(kFakeIO + align_up(in_bytes, 256)), in_bytes); trace_append(b);
    }
    run.forward(kFakeIO, kFakeIO + align_up(in_bytes, 256), key.u8 != 0, B, H, W, key.ch, key.cw, (float*)(kFakeIO + 2 * align_up(in_bytes, 256)));
  } catch (...) {
    s2m2b200_set_plan_mode(prev);
    if (pl.consts_on_device) cudaFree(pl.consts);
    else free(pl.consts);
    throw;
  }
  s2m2b200_set_plan_mode(prev);
  pl.workspace = run.arena.peak;
  return m.plans.emplace(key, pl).first->second;
}

}  // namespace host
}  // namespace s2m2

// ================================================================================================
// C-ABI
// ================================================================================================
#define S2M2_GUARD_BEGIN try {
#define S2M2_GUARD_END(errval)                     \
  }                                                \
  catch (const std::exception& e) {                \
    s2m2::set_error("%s", e.what());               \
    return errval;                                 \
  }

extern "C" {

int s2m2b200_create(int32_t feature_channels, int32_t num_transformer, int32_t use_positivity, int32_t output_upsample, int32_t refine_iter,
                    int32_t precision, S2m2Model** out) {
  S2M2_CHECK(out != nullptr, "create: out is NULL");
  S2M2_CHECK(feature_channels > 0 && feature_channels % 32 == 0 && num_transformer >= 1, "create: feature_channels must be a positive multiple of 32, num_transformer >= 1");
  S2M2_CHECK(precision == S2M2_F16 || precision == S2M2_F32, "create: precision must be S2M2_F16 or S2M2_F32");
  S2M2_CHECK(refine_iter >= 0, "create: refine_iter < 0");
  S2M2_GUARD_BEGIN
  S2m2Model* m = new S2m2Model();
  m->C = feature_channels; m->NTR = num_transformer; m->use_positivity = use_positivity ? 1 : 0; m->output_upsample = output_upsample ? 1 : 0;
  m->refine_iter = refine_iter; m->prec = precision;
  *out = m;
  return 0;
  S2M2_GUARD_END(-1)
}

void s2m2b200_destroy(S2m2Model* m) { delete m; }

int s2m2b200_set_param(S2m2Model* m, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  S2M2_CHECK(m && name && data && shape, "set_param: null argument");
  S2M2_CHECK(ndim >= 1 && ndim <= 4, "set_param(%s): ndim %d not in 1..4", name, ndim);
  S2M2_CHECK(!m->finalized, "set_param(%s): the parameters are already finalized", name);
  S2M2_GUARD_BEGIN
  HT t;
  t.nd = ndim;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    S2M2_CHECK(shape[i] > 0, "set_param(%s): bad shape", name);
    t.s[i] = shape[i];
    n *= shape[i];
  }
  t.v.assign(data, data + n);
  m->params[name] = std::move(t);
  return 0;
  S2M2_GUARD_END(-1)
}

int s2m2b200_finalize_params(S2m2Model* m) {
  S2M2_CHECK(m, "finalize_params: null model");
  S2M2_CHECK(!m->finalized, "finalize_params: already finalized");
  S2M2_GUARD_BEGIN
  Packer pk(*m);
  pk.pack_all();
  m->blob_bytes = align_up(m->staging.size(), 256);
  if (have_device()) {
    S2M2_CHECK(cudaMalloc((void**)&m->blob, m->blob_bytes) == cudaSuccess, "finalize_params: cudaMalloc(%zu) failed", m->blob_bytes);
    m->blob_on_device = true;
    S2M2_CHECK(cudaMemcpy(m->blob, m->staging.data(), m->staging.size(), cudaMemcpyHostToDevice) == cudaSuccess, "finalize_params: upload failed");
  } else {   // no device: the packed image stays on the host (inspection, device-less planning)
    m->blob = (char*)aligned_alloc(256, m->blob_bytes);
    S2M2_CHECK(m->blob != nullptr, "finalize_params: out of memory");
    memcpy(m->blob, m->staging.data(), m->staging.size());
  }
  std::vector<uint8_t>().swap(m->staging);
  std::map<std::string, HT>().swap(m->params);
  resolve(*m);
  m->finalized = true;
  return 0;
  S2M2_GUARD_END(-1)
}

int s2m2b200_set_graph_mode(S2m2Model* m, int32_t on) {
  S2M2_CHECK(m, "set_graph_mode: null model");
  m->graph_mode = on != 0;
  return 0;
}

int64_t s2m2b200_packed_count(const S2m2Model* m) { return m && m->finalized ? (int64_t)m->dir.size() : -1; }

int s2m2b200_packed_entry(const S2m2Model* m, int64_t index, const char** name, const void** ptr, int64_t* bytes, int32_t* dtype, int32_t* on_device) {
  S2M2_CHECK(m && m->finalized && index >= 0 && index < (int64_t)m->dir.size(), "packed_entry: bad model / index");
  const Entry& e = m->dir[(size_t)index];
  if (name) *name = e.name.c_str();
  if (ptr) *ptr = m->blob + e.off;
  if (bytes) *bytes = (int64_t)e.bytes;
  if (dtype) *dtype = e.dtype;
  if (on_device) *on_device = m->blob_on_device ? 1 : 0;
  return 0;
}

int64_t s2m2b200_workspace_bytes(S2m2Model* m, int32_t B, int32_t H, int32_t W, int32_t in_u8, int32_t crop_h, int32_t crop_w) {
  S2M2_CHECK(m && m->finalized, "workspace_bytes: model not finalized");
  S2M2_CHECK(B > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "workspace_bytes: H and W must be positive multiples of 32 (run_stereo_matching pads, model_utils.py:72-73)");
  S2M2_GUARD_BEGIN
  if (trace_on()) {
    char b[200];
    for (const Entry& e : m->dir) {
      snprintf(b, sizeof(b), "static %s %p %zu", e.name.c_str(), (void*)(m->blob + e.off), e.bytes);
      trace_append(b);
    }
    const struct { const char* n; const std::vector<float>* v; } small[] = {{"bb.w0", &m->bb_w0}, {"bb.b0", &m->bb_b0}, {"bb.w1", &m->bb_w1}, {"bb.b1", &m->bb_b1},
                                                                             {"u1.wd", &m->u1_wd}, {"u1.bd", &m->u1_bd}, {"u1.wr", &m->u1_wr}, {"u1.br", &m->u1_br}};
    for (const auto& s : small) {
      snprintf(b, sizeof(b), "static %s %p %zu", s.n, (const void*)s.v->data(), s.v->size() * 4);
      trace_append(b);
    }
  }
  const ShapeKey key{B, H, W, in_u8 ? 1 : 0, crop_h, crop_w};
  return (int64_t)make_plan(*m, key).workspace;
  S2M2_GUARD_END(-1)
}

int s2m2b200_forward(S2m2Model* m, const void* left, const void* right, int32_t in_u8, int32_t B, int32_t H, int32_t W, int32_t crop_h, int32_t crop_w,
                     float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  S2M2_CHECK(m && m->finalized, "forward: model not finalized");
  S2M2_CHECK(m->blob_on_device, "forward: the parameters were finalized without a CUDA device (there is no CPU path)");
  S2M2_CHECK(left && right && out && workspace, "forward: null pointer");
  S2M2_CHECK(B > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "forward: H and W must be positive multiples of 32 (run_stereo_matching pads, model_utils.py:72-73)");
  S2M2_CHECK(((uintptr_t)workspace & 255) == 0, "forward: the workspace must be 256-byte aligned");
  S2M2_GUARD_BEGIN
  const ShapeKey key{B, H, W, in_u8 ? 1 : 0, crop_h, crop_w};
  Plan& pl = make_plan(*m, key);   // cached after s2m2b200_workspace_bytes; first use of a shape plans here (not capturable)
  S2M2_CHECK((size_t)workspace_bytes >= pl.workspace, "forward: workspace of %lld bytes, %zu needed (s2m2b200_workspace_bytes)", (long long)workspace_bytes, pl.workspace);
  cudaStream_t cs = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool own_graph = m->graph_mode && cudaStreamIsCapturing(cs, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
  if (own_graph && pl.exec && pl.g_left == left && pl.g_right == right && pl.g_out == out && pl.g_work == workspace) {
    S2M2_CHECK(cudaGraphLaunch(pl.exec, cs) == cudaSuccess, "forward: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (own_graph && !m->cap_stream)
    S2M2_CHECK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking) == cudaSuccess, "forward: cudaStreamCreate failed");
  Run run(*m, own_graph ? (void*)m->cap_stream : stream, false);
  run.pl = &pl;
  run.arena.base = (char*)workspace;
  run.arena.cap = (size_t)workspace_bytes;
  if (!own_graph) {
    run.forward(left, right, in_u8 != 0, B, H, W, crop_h, crop_w, out);
    return 0;
  }
  // graph mode: capture this shape's ~900 launches once for these buffers, then replay (the addresses are baked into the graph;
  // a call with other buffers re-captures)
  if (pl.exec) {
    cudaGraphExecDestroy(pl.exec);
    pl.exec = nullptr;
  }
  S2M2_CHECK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess, "forward: cudaStreamBeginCapture failed: %s",
             cudaGetErrorString(cudaGetLastError()));
  cudaGraph_t graph = nullptr;
  try {
    run.forward(left, right, in_u8 != 0, B, H, W, crop_h, crop_w, out);
  } catch (...) {
    cudaStreamEndCapture(m->cap_stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  S2M2_CHECK(cudaStreamEndCapture(m->cap_stream, &graph) == cudaSuccess && graph, "forward: cudaStreamEndCapture failed: %s", cudaGetErrorString(cudaGetLastError()));
  const cudaError_t ie = cudaGraphInstantiate(&pl.exec, graph, 0);
  cudaGraphDestroy(graph);
  S2M2_CHECK(ie == cudaSuccess, "forward: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
  pl.g_left = left; pl.g_right = right; pl.g_out = out; pl.g_work = workspace;
  S2M2_CHECK(cudaGraphLaunch(pl.exec, cs) == cudaSuccess, "forward: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
  S2M2_GUARD_END(-1)
}

// ---- run_stereo_matching (model_utils.py:50-95) behind the handle: image_pad -> forward -> image_crop -> interior confidence mean ----
namespace {
struct MatchLayout {
  int Hn, Wn;
  size_t lp, rp, padwork, fwd, total;
};
MatchLayout match_layout(int B, int H0, int W0, size_t fwd_bytes) {
  MatchLayout l;
  l.Hn = (H0 + 31) / 32 * 32;
  l.Wn = (W0 + 31) / 32 * 32;
  const size_t img = align_up((size_t)B * 3 * l.Hn * l.Wn * 4, 256);
  l.lp = 0;
  l.rp = img;
  l.padwork = 2 * img;
  l.fwd = l.padwork + align_up((size_t)B * 3 * (H0 / 32) * (W0 / 32) * 4 + 4, 256);
  l.total = l.fwd + fwd_bytes;
  return l;
}
}  // namespace

int64_t s2m2b200_stereo_match_workspace_bytes(S2m2Model* m, int32_t B, int32_t H0, int32_t W0) {
  S2M2_CHECK(m && m->finalized, "stereo_match_workspace_bytes: model not finalized");
  S2M2_CHECK(B > 0 && H0 >= 32 && W0 >= 32, "stereo_match_workspace_bytes: images must be at least 32 x 32");
  S2M2_CHECK(!m->output_upsample, "stereo_match: not defined for output_upsample models (image_crop would cut the 2x maps)");
  const int Hn = (H0 + 31) / 32 * 32, Wn = (W0 + 31) / 32 * 32;
  const int64_t fwd = s2m2b200_workspace_bytes(m, B, Hn, Wn, 0, H0, W0);
  if (fwd < 0) return fwd;
  return (int64_t)match_layout(B, H0, W0, (size_t)fwd).total;
}

int s2m2b200_stereo_match(S2m2Model* m, const void* left, const void* right, int32_t in_u8, int32_t B, int32_t H0, int32_t W0, int32_t margin,
                          float* out, float* score, void* workspace, int64_t workspace_bytes, void* stream) {
  S2M2_CHECK(m && m->finalized && left && right && out && workspace, "stereo_match: null argument / model not finalized");
  S2M2_CHECK(B > 0 && H0 >= 32 && W0 >= 32, "stereo_match: images must be at least 32 x 32");
  S2M2_CHECK(!m->output_upsample, "stereo_match: not defined for output_upsample models");
  S2M2_CHECK(((uintptr_t)workspace & 255) == 0, "stereo_match: the workspace must be 256-byte aligned");
  const int Hn = (H0 + 31) / 32 * 32, Wn = (W0 + 31) / 32 * 32;
  const int64_t fwd = s2m2b200_workspace_bytes(m, B, Hn, Wn, 0, H0, W0);
  if (fwd < 0) return -1;
  const MatchLayout l = match_layout(B, H0, W0, (size_t)fwd);
  S2M2_CHECK((size_t)workspace_bytes >= l.total, "stereo_match: workspace of %lld bytes, %zu needed (s2m2b200_stereo_match_workspace_bytes)",
             (long long)workspace_bytes, l.total);
  char* w = (char*)workspace;
  float *lp = (float*)(w + l.lp), *rp = (float*)(w + l.rp), *pw = (float*)(w + l.padwork);
  // model_utils.py:72-73: both images padded to multiples of 32 with the blurred fill (image_utils.py:27-63)
  if (s2m2b200_image_pad(left, in_u8, B, H0, W0, 32, lp, pw, stream) != 0) return -1;
  if (s2m2b200_image_pad(right, in_u8, B, H0, W0, 32, rp, pw, stream) != 0) return -1;
  // the forward; its last kernel writes only the centre H0 x W0 window = image_crop (model_utils.py:88-91)
  if (s2m2b200_forward(m, lp, rp, 0, B, Hn, Wn, H0, W0, out, w + l.fwd, fwd, stream) != 0) return -1;
  if (score) {   // model_utils.py:93-94: mean of the cropped confidence map without a `margin`-wide frame
    if (margin > 0 && H0 > 2 * margin && W0 > 2 * margin) {
      if (s2m2b200_confidence_score(out + (size_t)2 * B * H0 * W0, B, H0, W0, H0, W0, margin, score, stream) != 0) return -1;
    } else {
      S2M2_CHECK(cudaMemsetAsync(score, 0xFF, (size_t)B * 4, (cudaStream_t)stream) == cudaSuccess, "stereo_match: memset failed");   // NaN: no interior
    }
  }
  return 0;
}

}  // extern "C"
