// CUDA-core implicit-GEMM convolution (fp32 FMA, fp32 accumulate) — the parity-mode GEMM engine and the
// fallback for shapes the tcgen05 kernel does not take (tiny channel counts, fp32 operands).
//
//   y[m, n] = act( sum_k A[m, k] * Wt[n, k] + bias[n] ) (+ res[m, n])
//   m = output pixel, k = (tap, channel) of the (optionally two-source) NHWC input, n = output channel.
//
// Reference ops replaced: see include/s2m2b200.h (S2m2Conv).
#include "common.cuh"

namespace s2m2 {

struct ConvP {
  const void *x, *x2, *w, *res;
  const float* bias;
  void* y;
  int xh, wh, rh, yh;  // fp16 flags
  long long x_ld, x2_ld, res_ld, y_ld;
  int H, W, C1, C2, Cin;
  int KH, KW, stride, ph, pw;
  int Ho, Wo, Cout, M, K;
  int act, shuffle, Cr, vout;
  long long x_bs, w_bs, y_bs;
};

template <int BM, int BN, int BK, int TM, int TN, bool VEC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), 2) conv_simt_kernel(ConvP p) {
  pdl_sync();
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int RM = TM / 4, RN = TN / 4;  // 4-wide register groups, spread over the tile to avoid bank conflicts
  constexpr int KV = BK / 4;
  static_assert(TM % 4 == 0 && TN % 4 == 0 && BK % 4 == 0, "tile shape");
  static_assert(NT % BM == 0 || VEC, "scalar path needs a fixed row per thread");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const long long zb = blockIdx.z;
  const char* xb = (const char*)p.x + zb * p.x_bs * (p.xh ? 2 : 4);
  const char* x2b = p.x2 ? (const char*)p.x2 + zb * p.x_bs * (p.xh ? 2 : 4) : nullptr;
  const char* wb = (const char*)p.w + zb * p.w_bs * (p.wh ? 2 : 4);
  char* yb = (char*)p.y + zb * p.y_bs * (p.yh ? 2 : 4);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // ---- per-thread load bookkeeping ----
  constexpr int LA = VEC ? (BM * KV + NT - 1) / NT : (BM * BK) / NT;
  constexpr int LB = VEC ? (BN * KV + NT - 1) / NT : (BN * BK + NT - 1) / NT;
  // pixel decode of the rows this thread loads
  int a_iy0[VEC ? LA : 1], a_ix0[VEC ? LA : 1];
  long long a_nb[VEC ? LA : 1];
  bool a_ok[VEC ? LA : 1];
  auto decode = [&](int m, int& iy0, int& ix0, long long& nb, bool& ok) {
    ok = m < p.M;
    int mm = ok ? m : 0;
    int ox = mm % p.Wo;
    int t = mm / p.Wo;
    int oy = t % p.Ho;
    int n = t / p.Ho;
    iy0 = oy * p.stride - p.ph;
    ix0 = ox * p.stride - p.pw;
    nb = (long long)n * p.H * p.W;
  };
  if (VEC) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      int f = tid + l * NT;
      decode(m0 + f / KV, a_iy0[l], a_ix0[l], a_nb[l], a_ok[l]);
      if (f >= BM * KV) a_ok[l] = false;
    }
  } else {
    decode(m0 + tid % BM, a_iy0[0], a_ix0[0], a_nb[0], a_ok[0]);
  }

  float4 a_reg[VEC ? LA : 1], b_reg[VEC ? LB : 1];
  float a_s[VEC ? 1 : LA], b_s[VEC ? 1 : LB];

  auto fetch = [&](int kt) {
    if (VEC) {
#pragma unroll
      for (int l = 0; l < LA; ++l) {
        int f = tid + l * NT;
        int k0 = kt * BK + (f % KV) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_ok[l] && k0 < p.K) {
          int tap = k0 / p.Cin, ci = k0 - tap * p.Cin;
          int ky = tap / p.KW, kx = tap - ky * p.KW;
          int iy = a_iy0[l] + ky, ix = a_ix0[l] + kx;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            long long pix = a_nb[l] + (long long)iy * p.W + ix;
            if (ci < p.C1) v = ld4(xb, pix * p.x_ld + ci, p.xh);
            else v = ld4(x2b, pix * p.x2_ld + (ci - p.C1), p.xh);
          }
        }
        a_reg[l] = v;
      }
#pragma unroll
      for (int l = 0; l < LB; ++l) {
        int f = tid + l * NT;
        int n = n0 + f / KV, k0 = kt * BK + (f % KV) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < BN * KV && n < p.Cout && k0 < p.K) v = ld4(wb, (long long)n * p.K + k0, p.wh);
        b_reg[l] = v;
      }
    } else {
#pragma unroll
      for (int l = 0; l < LA; ++l) {
        int k = kt * BK + (tid + l * NT) / BM;
        float v = 0.f;
        if (a_ok[0] && k < p.K) {
          int tap = k / p.Cin, ci = k - tap * p.Cin;
          int ky = tap / p.KW, kx = tap - ky * p.KW;
          int iy = a_iy0[0] + ky, ix = a_ix0[0] + kx;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            long long pix = a_nb[0] + (long long)iy * p.W + ix;
            if (ci < p.C1) v = ld1(xb, pix * p.x_ld + ci, p.xh);
            else v = ld1(x2b, pix * p.x2_ld + (ci - p.C1), p.xh);
          }
        }
        a_s[l] = v;
      }
#pragma unroll
      for (int l = 0; l < LB; ++l) {
        int e = tid + l * NT;
        int n = n0 + e % BN, k = kt * BK + e / BN;
        float v = 0.f;
        if (e < BN * BK && n < p.Cout && k < p.K) v = ld1(wb, (long long)n * p.K + k, p.wh);
        b_s[l] = v;
      }
    }
  };
  auto stash = [&]() {
    if (VEC) {
#pragma unroll
      for (int l = 0; l < LA; ++l) {
        int f = tid + l * NT;
        if (f < BM * KV) {
          int m = f / KV, k = (f % KV) * 4;
          As[k + 0][m] = a_reg[l].x; As[k + 1][m] = a_reg[l].y; As[k + 2][m] = a_reg[l].z; As[k + 3][m] = a_reg[l].w;
        }
      }
#pragma unroll
      for (int l = 0; l < LB; ++l) {
        int f = tid + l * NT;
        if (f < BN * KV) {
          int n = f / KV, k = (f % KV) * 4;
          Bs[k + 0][n] = b_reg[l].x; Bs[k + 1][n] = b_reg[l].y; Bs[k + 2][n] = b_reg[l].z; Bs[k + 3][n] = b_reg[l].w;
        }
      }
    } else {
#pragma unroll
      for (int l = 0; l < LA; ++l) As[(tid + l * NT) / BM][tid % BM] = a_s[l];
#pragma unroll
      for (int l = 0; l < LB; ++l) {
        int e = tid + l * NT;
        if (e < BN * BK) Bs[e / BN][e % BN] = b_s[l];
      }
    }
  };

  const int nk = (p.K + BK - 1) / BK;
  fetch(0);
  stash();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) fetch(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < RM; ++g) {
        float4 v = *(const float4*)&As[k][g * (BM / RM) + ty * 4];
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < RN; ++g) {
        float4 v = *(const float4*)&Bs[k][g * (BN / RN) + tx * 4];
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (kt + 1 < nk) {
      stash();
      __syncthreads();
    }
  }

  // ---- epilogue: bias, activation, residual, (pixel-shuffle) store ----
  const bool vec_out = p.vout != 0;
#pragma unroll
  for (int gi = 0; gi < RM; ++gi) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + gi * (BM / RM) + ty * 4 + r;
      if (m >= p.M) continue;
      long long opix = m;
      int ox = 0, oy = 0, nimg = 0;
      if (p.shuffle) {
        ox = m % p.Wo;
        int t = m / p.Wo;
        oy = t % p.Ho;
        nimg = t / p.Ho;
      }
#pragma unroll
      for (int gj = 0; gj < RN; ++gj) {
        const int n = n0 + gj * (BN / RN) + tx * 4;
        if (n >= p.Cout) continue;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = acc[gi * 4 + r][gj * 4 + c];
          if (p.bias && n + c < p.Cout) t += p.bias[n + c];
          v[c] = act_apply(t, p.act);
        }
        if (!p.shuffle) {
          if (vec_out) {
            if (p.res) {
              float4 rr = ld4(p.res, opix * p.res_ld + n, p.rh);
              v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
            }
            st4(yb, opix * p.y_ld + n, p.yh, make_float4(v[0], v[1], v[2], v[3]));
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (n + c < p.Cout) {
                float t = v[c];
                if (p.res) t += ld1(p.res, opix * p.res_ld + n + c, p.rh);
                st1(yb, opix * p.y_ld + n + c, p.yh, t);
              }
          }
        } else {
          // ConvTranspose2d(k=2,s=2): column n = q*Cr + co, q = a*2+b -> y[nimg, 2oy+a, 2ox+b, co]
          if (vec_out) {
            int q = n / p.Cr, co = n - q * p.Cr;
            long long pix = ((long long)nimg * 2 * p.Ho + 2 * oy + (q >> 1)) * (2 * p.Wo) + 2 * ox + (q & 1);
            st4(yb, pix * p.y_ld + co, p.yh, make_float4(v[0], v[1], v[2], v[3]));
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (n + c < p.Cout) {
                int q = (n + c) / p.Cr, co = (n + c) - q * p.Cr;
                long long pix = ((long long)nimg * 2 * p.Ho + 2 * oy + (q >> 1)) * (2 * p.Wo) + 2 * ox + (q & 1);
                st1(yb, pix * p.y_ld + co, p.yh, v[c]);
              }
          }
        }
      }
    }
  }
}

// ---- tiny-K layers (K = KH*KW*Cin <= 32: 1->C, 2->C, 9->C encoders of the refiner) -------------------------------
// The tiled GEMM above wastes its K loop on them; here a thread gathers its pixel's K inputs once and produces 8
// output channels from weights held in shared memory; 8-channel groups of one pixel are adjacent threads (coalesced).
constexpr int SMALLK_MAX = 32;
// VEC: fp16 output with 16-byte aligned 8-channel groups -> a thread owns 8 CONSECUTIVE channels (two LDS.128 of weights per k,
// one 16-byte store, exp-based activation); otherwise channels g, g+NG, .. with scalar stores (any dtype / alignment).
template <bool VEC>
__global__ void __launch_bounds__(256) conv_smallk_kernel(ConvP p) {
  pdl_sync();
  extern __shared__ float sw[];                  // [K][Cout] weights (transposed), [Cout] bias, [K] tap table
  float* sb = sw + p.Cout * p.K;
  int* stab = (int*)(sb + p.Cout);               // per k: (dy + 128) << 16 | (dx + 128) << 8 | ci
  for (int i = threadIdx.x; i < p.Cout * p.K; i += blockDim.x) {
    const int co = i / p.K, k = i - co * p.K;
    sw[k * p.Cout + co] = ld1(p.w, i, p.wh);
  }
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sb[i] = p.bias ? p.bias[i] : 0.f;
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    const int tap = k / p.Cin, ci = k - tap * p.Cin;
    const int ky = tap / p.KW, kx = tap - ky * p.KW;
    stab[k] = ((ky - p.ph + 128) << 16) | ((kx - p.pw + 128) << 8) | ci;
  }
  __syncthreads();
  const int NG = p.Cout >> 3;                    // threads per pixel
  const long long total = (long long)p.M * NG;
  if constexpr (VEC) {
    // the hot variant (the refiner's 1 / 2 / 9-channel feature layers, 12 launches per forward): the channel group of a thread is
    // fixed (blockDim is a multiple of NG), pixels advance by a constant stride, taps come from a table of element offsets — no
    // 64-bit division and no per-tap address arithmetic in the loop; GELU through the 10-instruction gelu_fast of the tcgen05 epilogues
    int* soff = stab + p.K;                      // per k: element offset (dy * W + dx) * x_ld + ci relative to the centre pixel
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
      const int e = stab[k];
      soff[k] = (((e >> 16) - 128) * p.W + (((e >> 8) & 255) - 128)) * (int)p.x_ld + (e & 255);
    }
    __syncthreads();
    const int g = threadIdx.x % NG, cb = 8 * g;
    const int ppb = blockDim.x / NG;             // pixels per block and iteration
    const long long mstep = (long long)gridDim.x * ppb;
    float4 b0 = *(const float4*)(sb + cb), b1 = *(const float4*)(sb + cb + 4);
    for (long long m = (long long)blockIdx.x * ppb + threadIdx.x / NG; m < p.M; m += mstep) {
      const int mi = (int)m;                     // M < 2^31 (checked by the launcher)
      const int ox = mi % p.Wo, t = mi / p.Wo;
      const int oy = t % p.Ho, n = t / p.Ho;
      const int cy = oy * p.stride, cx = ox * p.stride;
      const long long base = (((long long)n * p.H + cy) * p.W + cx) * p.x_ld;
      float o[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      for (int k = 0; k < p.K; ++k) {
        const int e = stab[k];
        const int iy = cy + (e >> 16) - 128, ix = cx + ((e >> 8) & 255) - 128;
        float a = 0.f;
        if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) a = ld1(p.x, base + soff[k], p.xh);
        const float4 w0 = *(const float4*)(sw + k * p.Cout + cb), w1 = *(const float4*)(sw + k * p.Cout + cb + 4);
        o[0] = fmaf(a, w0.x, o[0]); o[1] = fmaf(a, w0.y, o[1]); o[2] = fmaf(a, w0.z, o[2]); o[3] = fmaf(a, w0.w, o[3]);
        o[4] = fmaf(a, w1.x, o[4]); o[5] = fmaf(a, w1.y, o[5]); o[6] = fmaf(a, w1.z, o[6]); o[7] = fmaf(a, w1.w, o[7]);
      }
      __half2 h[4];
      if (p.act == S2M2_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(gelu_fast(o[2 * j]), gelu_fast(o[2 * j + 1]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(act_apply_fast(o[2 * j], p.act), act_apply_fast(o[2 * j + 1], p.act));
      }
      *(uint4*)((__half*)p.y + m * p.y_ld + cb) = *(uint4*)h;
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / NG), g = (int)(i - (long long)m * NG);
    const int ox = m % p.Wo;
    const int t = m / p.Wo;
    const int oy = t % p.Ho, n = t / p.Ho;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = sb[g + j * NG];
    const long long img = (long long)n * p.H;
    for (int k = 0; k < p.K; ++k) {
      const int e = stab[k];
      const int iy = oy * p.stride + (e >> 16) - 128, ix = ox * p.stride + ((e >> 8) & 255) - 128;
      float a = 0.f;
      if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) a = ld1(p.x, ((img + iy) * p.W + ix) * p.x_ld + (e & 255), p.xh);
      const float* wr = sw + k * p.Cout + g;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(a, wr[j * NG], o[j]);
    }
    const long long off = (long long)m * p.y_ld + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) st1(p.y, off + j * NG, p.yh, act_apply(o[j], p.act));
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_cfg(const ConvP& p, int batch, bool vec, cudaStream_t s) {
  dim3 grid(cdiv(p.M, BM), cdiv(p.Cout, BN), batch);
  dim3 block((BM / TM) * (BN / TN));
  if (vec) launch_k(conv_simt_kernel<BM, BN, BK, TM, TN, true>, grid, block, 0, s, p);
  else launch_k(conv_simt_kernel<BM, BN, BK, TM, TN, false>, grid, block, 0, s, p);
  S2M2_LAUNCH_OK("conv_simt");
  return 0;
}

int conv_simt_launch(const S2m2Conv* a, cudaStream_t s) {
  ConvP p{};
  const S2m2Tensor &x = a->x, &x2 = a->x2, &y = a->y, &res = a->res;
  p.x = x.ptr; p.x2 = x2.ptr; p.w = a->w; p.res = res.ptr; p.bias = a->bias; p.y = y.ptr;
  p.xh = x.dtype == S2M2_F16; p.wh = a->w_dtype == S2M2_F16; p.rh = res.ptr ? res.dtype == S2M2_F16 : 0;
  p.yh = y.dtype == S2M2_F16;
  p.x_ld = x.ld; p.x2_ld = x2.ptr ? x2.ld : 0; p.res_ld = res.ptr ? res.ld : 0; p.y_ld = y.ld;
  p.H = x.H; p.W = x.W; p.C1 = x.C; p.C2 = x2.ptr ? x2.C : 0; p.Cin = p.C1 + p.C2;
  p.KH = a->KH; p.KW = a->KW; p.stride = a->stride; p.ph = a->pad_h; p.pw = a->pad_w;
  p.Ho = (x.H + 2 * p.ph - p.KH) / p.stride + 1;
  p.Wo = (x.W + 2 * p.pw - p.KW) / p.stride + 1;
  p.Cout = a->Cout; p.M = x.N * p.Ho * p.Wo; p.K = p.KH * p.KW * p.Cin;
  p.act = a->act; p.shuffle = a->shuffle2x2; p.Cr = a->shuffle2x2 ? a->Cout / 4 : a->Cout;
  p.x_bs = a->x_bs; p.w_bs = a->w_bs; p.y_bs = a->y_bs;
  const int batch = a->batch > 0 ? a->batch : 1;

  bool vec = p.C1 % 4 == 0 && p.C2 % 4 == 0 && x.ld % 4 == 0 && aligned(x.ptr, 4 * esize(x.dtype)) &&
             aligned(a->w, 4 * esize(a->w_dtype)) && (p.x_bs % 4 == 0) && (p.w_bs % 4 == 0);
  if (x2.ptr) vec = vec && x2.ld % 4 == 0 && aligned(x2.ptr, 4 * esize(x2.dtype));
  // the vector epilogue additionally needs aligned bases (checked here, ld/Cout multiples checked in-kernel)
  bool out_al = aligned(y.ptr, 4 * esize(y.dtype)) && (p.y_bs % 4 == 0) && (!res.ptr || aligned(res.ptr, 4 * esize(res.dtype)));
  p.vout = out_al && p.y_ld % 4 == 0 &&
           (p.shuffle ? p.Cr % 4 == 0 : (p.Cout % 4 == 0 && (!res.ptr || res.ld % 4 == 0)));
  if (p.K <= SMALLK_MAX && p.Cout >= 16 && p.Cout % 8 == 0 && p.Cout <= 256 && batch == 1 && !p.x2 && !p.res && !p.shuffle) {
    const long long total = (long long)p.M * (p.Cout / 8);
    long long g = (total + 255) / 256;
    if (g > 148LL * 32) g = 148LL * 32;
    const size_t smem = sizeof(float) * (size_t)(p.Cout * p.K + p.Cout + 2 * p.K);
    // (the vector variant: fixed channel group per thread -> 256 % (Cout / 8) == 0; 32-bit tap offsets and pixel indices)
    const bool v8 = p.yh && p.y_ld % 8 == 0 && aligned(y.ptr, 16) && p.Cin < 256 && p.KH <= 128 && p.KW <= 128 && 256 % (p.Cout / 8) == 0 &&
                    p.Cout % 8 == 0 && (long long)x.N * p.H * p.W * p.x_ld < (1LL << 31) && p.M < (1LL << 31) && (long long)(p.KH + 1) * p.W * p.x_ld < (1LL << 30);
    if (v8) launch_k(conv_smallk_kernel<true>, (int)g, 256, smem, s, p);
    else launch_k(conv_smallk_kernel<false>, (int)g, 256, smem, s, p);
    S2M2_LAUNCH_OK("conv_smallk");
    return 0;
  }
  if (p.Cout > 64) return launch_cfg<128, 128, 16, 8, 8>(p, batch, vec, s);
  if (p.Cout > 32) return launch_cfg<128, 64, 16, 8, 4>(p, batch, vec, s);
  if (p.Cout > 16) return launch_cfg<128, 32, 16, 4, 4>(p, batch, vec, s);
  return launch_cfg<128, 16, 16, 4, 4>(p, batch, vec, s);
}

}  // namespace s2m2
