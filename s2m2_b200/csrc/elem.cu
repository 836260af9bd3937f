// Bandwidth-bound kernels: norms, resampling, gates, layout conversion, refiner glue.  All NHWC, any pixel
// stride, fp32 or fp16 storage (runtime flag), fp32 math.  Reference lines: see include/s2m2b200.h.
#include <string.h>

#include "common.cuh"

namespace s2m2 {

struct TV {  // device-side tensor view
  void* p;
  long long ld;
  int half;
  int C;
};
static inline TV tv(const S2m2Tensor& t) { return TV{t.ptr, t.ld, t.dtype == S2M2_F16, t.C}; }

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 32;  // a few waves of the 148 SMs; kernels are grid-stride
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over C (one warp per pixel, two-pass in registers)
// ------------------------------------------------------------------------------------------------
template <int MAXI>
__global__ void layernorm_kernel(TV x, TV y, const float* __restrict__ gamma, const float* __restrict__ beta, long long npix) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int C = x.C;
  for (long long pix = warp; pix < npix; pix += nwarps) {
    float v[MAXI];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      int c = lane + 32 * i;
      v[i] = c < C ? ld1(x.p, pix * x.ld + c, x.half) : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      int c = lane + 32 * i;
      float d = c < C ? v[i] - mean : 0.f;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float o = (v[i] - mean) * rstd;
        if (gamma) o = o * gamma[c] + beta[c];
        st1(y.p, pix * y.ld + c, y.half, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm: stats (sum, sumsq per (image, group) in double) then apply
// ------------------------------------------------------------------------------------------------
// Deterministic (atomic-free) reduction: per-block partial sums in a fixed order, then a fixed-order final sum.
// part: [N][nblk][groups][2] doubles, stats: [N][groups][2] doubles.
__global__ void gn_stats_kernel(TV x, int HW, int groups, int pix_per_block, double* __restrict__ part) {
  pdl_sync();
  extern __shared__ float sh[];  // [nw][C][2]
  const int n = blockIdx.y, C = x.C, cpg = C / groups;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = lane * 4; c < C; c += 128) {   // 4 channels per lane: 8/16-byte loads, a warp covers 128 channels of a pixel
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pp = p0 + wid; pp < p1; pp += nw) {
      const float4 v = ld4(x.p, ((long long)n * HW + pp) * x.ld + c, x.half);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sh[(wid * C + c + k) * 2 + 0] = s[k];
      sh[(wid * C + c + k) * 2 + 1] = q[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * groups) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    double acc = 0.0;
    for (int w = 0; w < nw; ++w)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) acc += (double)sh[(w * C + c) * 2 + k];
    part[(((long long)n * gridDim.x + blockIdx.x) * groups + g) * 2 + k] = acc;
  }
}

// one warp per (n, g, k): lanes stride over the block partials, then a fixed shuffle tree -> deterministic
// stats out: float2 (mean, rstd) per (n, g), computed once in double
__global__ void gn_final_kernel(const double* __restrict__ part, int nblk, int groups, int N, double cnt, float2* __restrict__ stats) {
  pdl_sync();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;  // (n, g)
  if (i >= N * groups) return;
  const int g = i % groups, n = i / groups;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += part[(((long long)n * nblk + b) * groups + g) * 2 + 0];
    q += part[(((long long)n * nblk + b) * groups + g) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    stats[i] = make_float2((float)mean, rsqrtf((float)var + 1e-5f));
  }
}

__global__ void gn_apply_kernel(TV x, TV y, int HW, int groups, const float2* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, long long total4) {
  pdl_sync();
  const int C = x.C, C4 = C / 4, cpg = C / groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i / C4;
    int c = (int)(i - pix * C4) * 4;
    int n = (int)(pix / HW);
    float4 v = ld4(x.p, pix * x.ld + c, x.half);
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 mr = __ldg(stats + (long long)n * groups + (c + k) / cpg);
      o[k] = (o[k] - mr.x) * mr.y * gamma[c + k] + beta[c + k];
    }
    st4(y.p, pix * y.ld + c, y.half, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ------------------------------------------------------------------------------------------------
// resampling
// ------------------------------------------------------------------------------------------------
__global__ void avgpool2_kernel(TV x, TV y, int N, int H, int W, long long total4) {
  pdl_sync();
  const int C4 = x.C / 4, Ho = H / 2, Wo = W / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i / C4;
    int c = (int)(i - pix * C4) * 4;
    int ox = (int)(pix % Wo);
    long long t = pix / Wo;
    int oy = (int)(t % Ho);
    long long n = t / Ho;
    long long b = (n * H + 2 * oy) * W + 2 * ox;
    float4 a0 = ld4(x.p, b * x.ld + c, x.half), a1 = ld4(x.p, (b + 1) * x.ld + c, x.half);
    float4 a2 = ld4(x.p, (b + W) * x.ld + c, x.half), a3 = ld4(x.p, (b + W + 1) * x.ld + c, x.half);
    float4 o = make_float4((a0.x + a1.x + a2.x + a3.x) * 0.25f, (a0.y + a1.y + a2.y + a3.y) * 0.25f,
                           (a0.z + a1.z + a2.z + a3.z) * 0.25f, (a0.w + a1.w + a2.w + a3.w) * 0.25f);
    st4(y.p, pix * y.ld + c, y.half, o);
  }
}

// bilinear x2, align_corners=False: src = (dst + .5)/2 - .5 clamped at 0 (torch semantics)
__global__ void upsample2x_kernel(TV x, TV y, int N, int H, int W, long long total4) {
  pdl_sync();
  const int C4 = x.C / 4, Ho = 2 * H, Wo = 2 * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i / C4;
    int c = (int)(i - pix * C4) * 4;
    int ox = (int)(pix % Wo);
    long long t = pix / Wo;
    int oy = (int)(t % Ho);
    long long n = t / Ho;
    float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    long long r0 = (n * H + y0) * W, r1 = (n * H + y1) * W;
    float4 a = ld4(x.p, (r0 + x0) * x.ld + c, x.half), b = ld4(x.p, (r0 + x1) * x.ld + c, x.half);
    float4 d = ld4(x.p, (r1 + x0) * x.ld + c, x.half), e = ld4(x.p, (r1 + x1) * x.ld + c, x.half);
    float4 o;
    o.x = hy * (hx * a.x + lx * b.x) + ly * (hx * d.x + lx * e.x);
    o.y = hy * (hx * a.y + lx * b.y) + ly * (hx * d.y + lx * e.y);
    o.z = hy * (hx * a.z + lx * b.z) + ly * (hx * d.z + lx * e.z);
    o.w = hy * (hx * a.w + lx * b.w) + ly * (hx * d.w + lx * e.w);
    st4(y.p, pix * y.ld + c, y.half, o);
  }
}


// ------------------------------------------------------------------------------------------------
// fp16 fast paths of the resampling / GroupNorm kernels: 8 channels (16 B) per thread, the channel slot is fixed per thread
// (blockDim = C/8 x pixels-per-block), rows are dealt to blocks — no integer division in the pixel loop, full-sector accesses.
// Same arithmetic as the generic kernels above.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __half2* h = (const __half2*)&v;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __half2* h = (__half2*)&v;
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// Bilinear x2 (align_corners = False: weights .75 / .25, clamped at the border).  One thread turns a 2x2 block of input pixels
// (columns i-1, i and rows j-1, j, clamped) into the 2x2 block of output pixels (2i-1, 2i) x (2j-1, 2j) that depends on exactly those
// four: 4 loads + 4 unpacks per FOUR outputs (the one-output-per-thread version spent its time unpacking: 29 % of the HBM peak).
__global__ void upsample2x_h8_kernel(const __half* __restrict__ x, long long xld, __half* __restrict__ y, long long yld, int N, int H, int W,
                                     int C8, int PPB) {
  pdl_sync();
  const int c8 = threadIdx.x % C8, p0 = threadIdx.x / C8;
  const int Wo = 2 * W;
  for (int row = blockIdx.x; row < N * (H + 1); row += gridDim.x) {
    const int n = row / (H + 1), j = row - n * (H + 1);
    const int ya = max(j - 1, 0), yb = min(j, H - 1);
    const __half* r0 = x + ((long long)n * H + ya) * W * xld + c8 * 8;
    const __half* r1 = x + ((long long)n * H + yb) * W * xld + c8 * 8;
    __half* o0 = y + ((long long)n * 2 * H + 2 * j - 1) * Wo * yld + c8 * 8;   // output row 2j-1 (exists for j >= 1)
    __half* o1 = o0 + (long long)Wo * yld;                                      // output row 2j   (exists for j <= H-1)
    const bool top = j >= 1, bot = j <= H - 1;
    for (int i = p0; i <= W; i += PPB) {
      const int xa = max(i - 1, 0), xb = min(i, W - 1);
      float a[8], b[8], d[8], e[8], t0[8], t1[8], u0[8], u1[8], o[8];
      unpack8(*(const uint4*)(r0 + xa * xld), a); unpack8(*(const uint4*)(r0 + xb * xld), b);
      unpack8(*(const uint4*)(r1 + xa * xld), d); unpack8(*(const uint4*)(r1 + xb * xld), e);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        t0[k] = 0.75f * a[k] + 0.25f * b[k]; t1[k] = 0.25f * a[k] + 0.75f * b[k];    // row ya at output columns 2i-1, 2i
        u0[k] = 0.75f * d[k] + 0.25f * e[k]; u1[k] = 0.25f * d[k] + 0.75f * e[k];    // row yb
      }
      const bool lft = i >= 1, rgt = i <= W - 1;
      if (top && lft) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.75f * t0[k] + 0.25f * u0[k];
        *(uint4*)(o0 + (long long)(2 * i - 1) * yld) = pack8(o);
      }
      if (top && rgt) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.75f * t1[k] + 0.25f * u1[k];
        *(uint4*)(o0 + (long long)(2 * i) * yld) = pack8(o);
      }
      if (bot && lft) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.25f * t0[k] + 0.75f * u0[k];
        *(uint4*)(o1 + (long long)(2 * i - 1) * yld) = pack8(o);
      }
      if (bot && rgt) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.25f * t1[k] + 0.75f * u1[k];
        *(uint4*)(o1 + (long long)(2 * i) * yld) = pack8(o);
      }
    }
  }
}

__global__ void avgpool2_h8_kernel(const __half* __restrict__ x, long long xld, __half* __restrict__ y, long long yld, int N, int H, int W,
                                   int C8, int PPB) {
  pdl_sync();
  const int c8 = threadIdx.x % C8, p0 = threadIdx.x / C8;
  const int Ho = H / 2, Wo = W / 2;
  for (int row = blockIdx.x; row < N * Ho; row += gridDim.x) {
    const int n = row / Ho, oy = row - n * Ho;
    const __half* r0 = x + ((long long)n * H + 2 * oy) * W * xld + c8 * 8;
    const __half* r1 = r0 + (long long)W * xld;
    __half* out = y + (long long)row * Wo * yld + c8 * 8;
    for (int ox = p0; ox < Wo; ox += PPB) {
      float a[8], b[8], d[8], e[8], o[8];
      unpack8(*(const uint4*)(r0 + (2 * ox) * xld), a); unpack8(*(const uint4*)(r0 + (2 * ox + 1) * xld), b);
      unpack8(*(const uint4*)(r1 + (2 * ox) * xld), d); unpack8(*(const uint4*)(r1 + (2 * ox + 1) * xld), e);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (a[k] + b[k] + d[k] + e[k]) * 0.25f;
      *(uint4*)(out + ox * yld) = pack8(o);
    }
  }
}

// GroupNorm statistics, fp16: block = (C/8) x PPB threads over a contiguous range of pixels of image blockIdx.y; partial sums per
// block in a fixed order (deterministic), same `part` layout as gn_stats_kernel.
__global__ void gn_stats_h8_kernel(const __half* __restrict__ x, long long xld, int HW, int C8, int PPB, int groups, int pix_per_block,
                                   double* __restrict__ part) {
  pdl_sync();
  extern __shared__ float sh[];  // [PPB][C][2]
  const int n = blockIdx.y, C = C8 * 8, cpg = C / groups;
  const int c8 = threadIdx.x % C8, p0 = threadIdx.x / C8;
  const int b0 = blockIdx.x * pix_per_block, b1 = min(b0 + pix_per_block, HW);
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  const __half* base = x + (long long)n * HW * xld + c8 * 8;
  for (int pp = b0 + p0; pp < b1; pp += PPB) {
    float v[8];
    unpack8(*(const uint4*)(base + pp * xld), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sh[(p0 * C + c8 * 8 + k) * 2 + 0] = s[k];
    sh[(p0 * C + c8 * 8 + k) * 2 + 1] = q[k];
  }
  __syncthreads();
  if (threadIdx.x < 2 * groups) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    double acc = 0.0;
    for (int w = 0; w < PPB; ++w)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) acc += (double)sh[(w * C + c) * 2 + k];
    part[(((long long)n * gridDim.x + blockIdx.x) * groups + g) * 2 + k] = acc;
  }
}

__global__ void gn_apply_h8_kernel(const __half* __restrict__ x, long long xld, __half* __restrict__ y, long long yld, int HW, int C8, int PPB,
                                   int groups, const float2* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta) {
  pdl_sync();
  const int n = blockIdx.y, C = C8 * 8, cpg = C / groups;
  const int c8 = threadIdx.x % C8, p0 = threadIdx.x / C8;
  float mean[8], sc[8], sf[8];   // same operation order as the generic kernel: (v - mean) * rstd * gamma + beta
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c8 * 8 + k;
    const float2 mr = __ldg(stats + (long long)n * groups + c / cpg);
    mean[k] = mr.x; sc[k] = mr.y; sf[k] = beta[c];
  }
  float gm[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gm[k] = gamma[c8 * 8 + k];
  const __half* xb = x + (long long)n * HW * xld + c8 * 8;
  __half* yb = y + (long long)n * HW * yld + c8 * 8;
  for (int pp = blockIdx.x * PPB + p0; pp < HW; pp += gridDim.x * PPB) {
    float v[8];
    unpack8(*(const uint4*)(xb + pp * xld), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean[k]) * sc[k] * gm[k] + sf[k];
    *(uint4*)(yb + pp * yld) = pack8(v);
  }
}

static inline bool h8_ok(const S2m2Tensor& t) {
  return t.dtype == S2M2_F16 && t.C % 8 == 0 && t.C <= 2048 && t.ld % 8 == 0 && aligned(t.ptr, 16);
}

// ------------------------------------------------------------------------------------------------
// element-wise (vector of 4 channels per thread)
// ------------------------------------------------------------------------------------------------
enum { EW_COPY = 0, EW_UNARY = 1, EW_MUL = 2, EW_GRU = 3, EW_BLEND = 4 };

template <int OP>
__global__ void ew_kernel(TV a, TV b, TV c, TV d, TV o, int act, long long total4) {
  pdl_sync();
  const int C4 = o.C / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i / C4;
    int ch = (int)(i - pix * C4) * 4;
    float4 r;
    if (OP == EW_COPY) {
      r = ld4(a.p, pix * a.ld + ch, a.half);
    } else if (OP == EW_UNARY) {
      float4 v = ld4(a.p, pix * a.ld + ch, a.half);
      r = make_float4(act_apply(v.x, act), act_apply(v.y, act), act_apply(v.z, act), act_apply(v.w, act));
    } else if (OP == EW_MUL) {
      float4 v = ld4(a.p, pix * a.ld + ch, a.half), w = ld4(b.p, pix * b.ld + ch, b.half);
      r = make_float4(v.x * w.x, v.y * w.y, v.z * w.z, v.w * w.w);
    } else if (OP == EW_GRU) {  // a=z, b=q, o=h (in place): h = (1-z)h + z q
      float4 z = ld4(a.p, pix * a.ld + ch, a.half), q = ld4(b.p, pix * b.ld + ch, b.half);
      float4 h = ld4(o.p, pix * o.ld + ch, o.half);
      r = make_float4((1.f - z.x) * h.x + z.x * q.x, (1.f - z.y) * h.y + z.y * q.y, (1.f - z.z) * h.z + z.z * q.z,
                      (1.f - z.w) * h.w + z.w * q.w);
    } else {  // EW_BLEND: a=f, b=g, c=z0, d=z1
      float4 f = ld4(a.p, pix * a.ld + ch, a.half), g = ld4(b.p, pix * b.ld + ch, b.half);
      float4 z0 = ld4(c.p, pix * c.ld + ch, c.half), z1 = ld4(d.p, pix * d.ld + ch, d.half);
      float w0 = fminf(fmaxf(g.x, 0.01f), 0.99f), w1 = fminf(fmaxf(g.y, 0.01f), 0.99f);
      float w2 = fminf(fmaxf(g.z, 0.01f), 0.99f), w3 = fminf(fmaxf(g.w, 0.01f), 0.99f);
      r = make_float4(f.x + w0 * z0.x + (1.f - w0) * z1.x, f.y + w1 * z0.y + (1.f - w1) * z1.y,
                      f.z + w2 * z0.z + (1.f - w2) * z1.z, f.w + w3 * z0.w + (1.f - w3) * z1.w);
    }
    st4(o.p, pix * o.ld + ch, o.half, r);
  }
}

// TIn = float (what image_pad hands to the model, model_utils.py:72-73) or uint8_t (4x fewer input bytes; the
// values are the same integers, so both instantiations produce bit-identical results)
template <typename TIn>
__global__ void preprocess_kernel(const TIn* __restrict__ left, const TIn* __restrict__ right, int B, int H, int W, TV o) {
  pdl_sync();
  const long long HW = (long long)H * W, total = 2LL * B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long n = i / HW, hw = i - n * HW;
    const TIn* src = n < B ? left + n * 3 * HW : right + (n - B) * 3 * HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) st1(o.p, i * o.ld + c, o.half, ((float)src[c * HW + hw] / 255.f - 0.5f) * 2.f);
  }
}


// ------------------------------------------------------------------------------------------------
// fused full-resolution tiny-channel layers (thread per pixel; weights in shared memory)
// ------------------------------------------------------------------------------------------------
// stem: normalise -> Conv1x1(3->16) -> GELU -> Conv1x1(16->16)   (s2m2.py:86-87, submodules.py:70-72)
struct StemW { float w0[48], b0[16], w1[256], b1[16]; };      // passed by value: constant bank, warp-uniform reads
template <typename TIn>
__global__ void __launch_bounds__(256) stem_kernel(const TIn* __restrict__ left, const TIn* __restrict__ right, int B, int H, int W,
                                                   const __grid_constant__ StemW wts, TV o, int s2d) {
  pdl_sync();
  const float* sw0 = wts.w0; const float* sb0 = wts.b0; const float* sw1 = wts.w1; const float* sb1 = wts.b1;
  const long long HW = (long long)H * W, total = 2LL * B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long n = i / HW, hw = i - n * HW;
    const TIn* src = n < B ? left + n * 3 * HW : right + (n - B) * 3 * HW;
    float x[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = ((float)src[c * HW + hw] / 255.f - 0.5f) * 2.f;
    float hdn[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float t = sb0[j] + sw0[j * 3] * x[0] + sw0[j * 3 + 1] * x[1] + sw0[j * 3 + 2] * x[2];
      hdn[j] = 0.5f * t * (1.0f + erff(t * 0.70710678118654752440f));
    }
    float out[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float t = sb1[k];
#pragma unroll
      for (int j = 0; j < 16; ++j) t = fmaf(sw1[k * 16 + j], hdn[j], t);
      out[k] = t;
    }
    if (s2d) {   // space-to-depth: out[n, y/2, x/2, ((y&1)*2 + (x&1))*16 + k]  (fp16, 32 contiguous bytes per pixel)
      const int yy = (int)(hw / W), xx = (int)(hw - (long long)yy * W);
      const long long q = (n * (H / 2) + (yy >> 1)) * (W / 2) + (xx >> 1);
      __half* dst = (__half*)o.p + q * o.ld + ((yy & 1) * 2 + (xx & 1)) * 16;
      *(uint4*)dst = pack8(out);
      *(uint4*)(dst + 8) = pack8(out + 8);
      continue;
    }
#pragma unroll
    for (int k = 0; k < 16; k += 4) st4(o.p, i * o.ld + k, o.half, make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]));
    for (int k = 16; k < o.C; k += 4) st4(o.p, i * o.ld + k, o.half, make_float4(0.f, 0.f, 0.f, 0.f));   // zero channel padding
  }
}

// UpsampleMask1x inputs: relu(convT3x3(disp_up, 1->16)) | relu(convT3x3(rgb, 3->16))  (submodules.py:121-124,138-139)
// weights are the equivalent (flipped) conv weights, tap-major: wd[co][tap], wr[co][tap*3 + ci]
struct Mask1xW { float wd[144], wr[432], bd[16], br[16]; };
__global__ void __launch_bounds__(256) mask1x_inputs_kernel(const float4* __restrict__ up, TV rgb, int B, int H, int W,
                                                            const __grid_constant__ Mask1xW wts, TV o) {
  pdl_sync();
  const float* swd = wts.wd; const float* swr = wts.wr; const float* sbd = wts.bd; const float* sbr = wts.br;
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % W);
    long long t = i / W;
    int Y = (int)(t % H);
    long long b = t / H;
    float dn[9], cn[27];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int tap = (dy + 1) * 3 + dx + 1;
        const int yy = Y + dy, xx = X + dx;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const long long q = (b * H + yy) * W + xx;
        dn[tap] = in ? up[q].x : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) cn[tap * 3 + c] = in ? ld1(rgb.p, q * rgb.ld + c, rgb.half) : 0.f;
      }
    float out[32];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a = sbd[k], r = sbr[k];
#pragma unroll
      for (int j = 0; j < 9; ++j) a = fmaf(swd[k * 9 + j], dn[j], a);
#pragma unroll
      for (int j = 0; j < 27; ++j) r = fmaf(swr[k * 27 + j], cn[j], r);
      out[k] = fmaxf(a, 0.f);
      out[16 + k] = fmaxf(r, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 32; k += 4) st4(o.p, i * o.ld + k, o.half, make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]));
  }
}

// ------------------------------------------------------------------------------------------------
// refiner glue on the 1/4-resolution scalar maps (fp32)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float logit_eps(float x, float eps) {
  x = fminf(fmaxf(x, eps), 1.f - eps);
  return logf(x / (1.f - x));
}

__global__ void gr_prep_kernel(const float* __restrict__ disp, const float* __restrict__ conf, long long n, TV o) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float c = conf[i], m = c > 0.2f ? 1.f : 0.f;
    st1(o.p, i * o.ld + 0, o.half, disp[i] / 100.f * m);
    st1(o.p, i * o.ld + 1, o.half, logit_eps(m * c, 0.1f));
  }
}

__global__ void gr_merge_kernel(float* __restrict__ disp, const float* __restrict__ conf, const float* __restrict__ upd,
                                long long upd_ld, long long n, int pos) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float m = conf[i] > 0.2f ? 1.f : 0.f;
    float d = m * disp[i] + (1.f - m) * upd[i * upd_ld];
    disp[i] = pos ? fmaxf(d, 0.f) : d;
  }
}

__global__ void lr_prep_kernel(const float* __restrict__ disp, const float* __restrict__ conf, const float* __restrict__ occ,
                               long long n, float4* __restrict__ small) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    small[i] = make_float4(disp[i] / 100.f, logit_eps(conf[i], 0.01f), logit_eps(occ[i], 0.01f), 0.f);
}

__global__ void lr_update_kernel(float* __restrict__ disp, float* __restrict__ conf, float* __restrict__ occ,
                                 const float4* __restrict__ small, const float* __restrict__ upd, long long upd_ld,
                                 long long n, int w, int pos) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float4 s = small[i];
    const float* u = upd + i * upd_ld;   // [delta disp, delta conf logit, delta occ logit, ...]
    float d = disp[i] + u[0];
    if (pos) d = fmaxf(d, 0.f);
    float x = (float)(i % w);
    conf[i] = 1.f / (1.f + expf(-(u[1] + s.y)));
    float o = 1.f / (1.f + expf(-(u[2] + s.z)));
    occ[i] = (x - d >= 0.f) ? o : 0.f;
    disp[i] = d;
  }
}

}  // namespace s2m2

using namespace s2m2;

// ==================================================================================================
// C-ABI
// ==================================================================================================
template <typename TIn>
static int preprocess_any(const TIn* left, const TIn* right, int32_t B, int32_t H, int32_t W, const S2m2Tensor* out, void* stream) {
  S2M2_CHECK(out->N == 2 * B && out->H == H && out->W == W && out->C == 3, "preprocess: output must be (2B,H,W,3)");
  launch_k(preprocess_kernel<TIn>, grid_for(2LL * B * H * W), 256, 0, (cudaStream_t)stream, left, right, B, H, W, tv(*out));
  S2M2_LAUNCH_OK("preprocess");
  return 0;
}

template <typename TIn>
static int stem_any(const TIn* left, const TIn* right, int32_t B, int32_t H, int32_t W, const float* w0, const float* b0,
                    const float* w1, const float* b1, const S2m2Tensor* out, void* stream) {
  // two output layouts: (2B,H,W,>=16) pixel-major, or (2B,H/2,W/2,64) fp16 space-to-depth (2x2 pixel blocks in the channel dim),
  // which turns the following 5x5 stride-2 convolution into a dense 3x3 stride-1 one over 64 channels
  const int s2d = (out->H == H / 2 && out->W == W / 2 && out->C == 64 && H % 2 == 0 && W % 2 == 0) ? 1 : 0;
  if (s2d) S2M2_CHECK(out->N == 2 * B && out->dtype == S2M2_F16 && out->ld % 8 == 0 && aligned(out->ptr, 16),
                      "stem: the space-to-depth output must be (2B,H/2,W/2,64) fp16 with 16-byte aligned pixels");
  else S2M2_CHECK(out->N == 2 * B && out->H == H && out->W == W && out->C >= 16 && vec4_ok(*out),
             "stem: output must be (2B,H,W,>=16), 4-aligned (channels beyond 16 are zero-filled)");
  StemW hw;   // w0/b0/w1/b1 are HOST pointers: the 1.3 KB of weights travel as a kernel parameter (constant bank)
  memcpy(hw.w0, w0, sizeof(hw.w0)); memcpy(hw.b0, b0, sizeof(hw.b0)); memcpy(hw.w1, w1, sizeof(hw.w1)); memcpy(hw.b1, b1, sizeof(hw.b1));
  launch_k(stem_kernel<TIn>, grid_for(2LL * B * H * W), 256, 0, (cudaStream_t)stream, left, right, B, H, W, hw, tv(*out), s2d);
  S2M2_LAUNCH_OK("stem");
  return 0;
}

extern "C" {

int s2m2b200_layernorm(const S2m2Tensor* x, const S2m2Tensor* y, const float* gamma, const float* beta, void* stream) {
  S2M2_TRACE("layernorm", x, y, gamma, beta);
  S2M2_CHECK(x->C == y->C && npix(*x) == npix(*y), "layernorm: shape mismatch");
  S2M2_CHECK(x->C <= 1024, "layernorm: C=%d > 1024", x->C);
  S2M2_CHECK((gamma == nullptr) == (beta == nullptr), "layernorm: gamma/beta must both be given or both NULL");
  long long n = npix(*x);
  int block = 256, grid = grid_for(n * 32, block);
  cudaStream_t s = (cudaStream_t)stream;
  if (x->C <= 128) launch_k(layernorm_kernel<4>, grid, block, 0, s, tv(*x), tv(*y), gamma, beta, n);
  else if (x->C <= 256) launch_k(layernorm_kernel<8>, grid, block, 0, s, tv(*x), tv(*y), gamma, beta, n);
  else if (x->C <= 512) launch_k(layernorm_kernel<16>, grid, block, 0, s, tv(*x), tv(*y), gamma, beta, n);
  else launch_k(layernorm_kernel<32>, grid, block, 0, s, tv(*x), tv(*y), gamma, beta, n);
  S2M2_LAUNCH_OK("layernorm");
  return 0;
}

int s2m2b200_groupnorm(const S2m2Tensor* x, const S2m2Tensor* y, int32_t groups, const float* gamma, const float* beta,
                       double* work, void* stream) {
  S2M2_TRACE("groupnorm", x, y, groups, gamma, beta, work);
  S2M2_CHECK(x->C == y->C && npix(*x) == npix(*y), "groupnorm: shape mismatch");
  S2M2_CHECK(groups > 0 && x->C % groups == 0 && groups <= 64, "groupnorm: bad groups");
  S2M2_CHECK(vec4_ok(*x) && vec4_ok(*y), "groupnorm: needs C,ld multiples of 4 and aligned bases");
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = x->H * x->W;
  const int ppb = 256, nblk = cdiv(HW, ppb);
  // work layout: stats [N][groups][2] doubles, then partials [N][nblk][groups][2] doubles
  double* stats = work;
  double* part = work + 2LL * x->N * groups;
  dim3 grid(nblk, x->N);
  const bool h8 = h8_ok(*x) && h8_ok(*y) && x->C <= 512;
  const int C8 = x->C / 8, PPB = h8 ? 256 / C8 : 8;
  const size_t smem = sizeof(float) * 2 * PPB * x->C;
  S2M2_CHECK(smem <= 48 * 1024, "groupnorm: C=%d too large", x->C);
  if (h8) launch_k(gn_stats_h8_kernel, grid, C8 * PPB, smem, s, (const __half*)x->ptr, x->ld, HW, C8, PPB, groups, ppb, part);
  else launch_k(gn_stats_kernel, grid, 256, smem, s, tv(*x), HW, groups, ppb, part);
  S2M2_LAUNCH_OK("gn_stats");
  launch_k(gn_final_kernel, cdiv((long long)x->N * groups * 32, 128), 128, 0, s, part, nblk, groups, x->N, (double)HW * (x->C / groups),
                                                                           (float2*)stats);
  S2M2_LAUNCH_OK("gn_final");
  if (h8) {
    const int bpi = cdiv(HW, PPB * 8);   // ~8 pixels per thread
    launch_k(gn_apply_h8_kernel, dim3(bpi < 148 * 8 ? bpi : 148 * 8, x->N), C8 * PPB, 0, s, (const __half*)x->ptr, x->ld, (__half*)y->ptr, y->ld,
             HW, C8, PPB, groups, (const float2*)stats, gamma, beta);
    S2M2_LAUNCH_OK("gn_apply_h8");
    return 0;
  }
  long long total4 = npix(*x) * (x->C / 4);
  launch_k(gn_apply_kernel, grid_for(total4), 256, 0, s, tv(*x), tv(*y), HW, groups, (const float2*)stats, gamma, beta, total4);
  S2M2_LAUNCH_OK("gn_apply");
  return 0;
}

int s2m2b200_avgpool2(const S2m2Tensor* x, const S2m2Tensor* y, void* stream) {
  S2M2_TRACE("avgpool2", x, y);
  S2M2_CHECK(x->C == y->C && y->H == x->H / 2 && y->W == x->W / 2 && x->N == y->N && x->H % 2 == 0 && x->W % 2 == 0,
             "avgpool2: shape mismatch");
  S2M2_CHECK(vec4_ok(*x) && vec4_ok(*y), "avgpool2: needs C,ld multiples of 4 and aligned bases");
  if (h8_ok(*x) && h8_ok(*y)) {
    const int C8 = x->C / 8, PPB = C8 >= 256 ? 1 : 256 / C8;
    const int rows = y->N * y->H;
    launch_k(avgpool2_h8_kernel, rows < 148 * 16 ? rows : 148 * 16, C8 * PPB, 0, (cudaStream_t)stream, (const __half*)x->ptr, x->ld,
             (__half*)y->ptr, y->ld, x->N, x->H, x->W, C8, PPB);
    S2M2_LAUNCH_OK("avgpool2_h8");
    return 0;
  }
  long long total4 = npix(*y) * (y->C / 4);
  launch_k(avgpool2_kernel, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*x), tv(*y), x->N, x->H, x->W, total4);
  S2M2_LAUNCH_OK("avgpool2");
  return 0;
}

int s2m2b200_upsample2x(const S2m2Tensor* x, const S2m2Tensor* y, void* stream) {
  S2M2_TRACE("upsample2x", x, y);
  S2M2_CHECK(x->C == y->C && y->H == x->H * 2 && y->W == x->W * 2 && x->N == y->N, "upsample2x: shape mismatch");
  S2M2_CHECK(vec4_ok(*x) && vec4_ok(*y), "upsample2x: needs C,ld multiples of 4 and aligned bases");
  if (h8_ok(*x) && h8_ok(*y)) {
    const int C8 = x->C / 8, PPB = C8 >= 256 ? 1 : 256 / C8;
    const int rows = x->N * (x->H + 1);      // one block iteration per pair of output rows (2j-1, 2j), j = 0..H
    launch_k(upsample2x_h8_kernel, rows < 148 * 16 ? rows : 148 * 16, C8 * PPB, 0, (cudaStream_t)stream, (const __half*)x->ptr, x->ld,
             (__half*)y->ptr, y->ld, x->N, x->H, x->W, C8, PPB);
    S2M2_LAUNCH_OK("upsample2x_h8");
    return 0;
  }
  long long total4 = npix(*y) * (y->C / 4);
  launch_k(upsample2x_kernel, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*x), tv(*y), x->N, x->H, x->W, total4);
  S2M2_LAUNCH_OK("upsample2x");
  return 0;
}

static int ew_check(const S2m2Tensor* o, const S2m2Tensor* const* ins, int n, const char* name) {
  S2M2_CHECK(vec4_ok(*o), "%s: output needs C,ld multiples of 4 and an aligned base", name);
  for (int i = 0; i < n; ++i) {
    S2M2_CHECK(ins[i]->C == o->C && npix(*ins[i]) == npix(*o), "%s: shape mismatch (input %d)", name, i);
    S2M2_CHECK(vec4_ok(*ins[i]), "%s: input %d needs C,ld multiples of 4 and an aligned base", name, i);
  }
  return 0;
}

int s2m2b200_gate_blend(const S2m2Tensor* f, const S2m2Tensor* g, const S2m2Tensor* z0, const S2m2Tensor* z1,
                        const S2m2Tensor* out, void* stream) {
  S2M2_TRACE("gate_blend", f, g, z0, z1, out);
  const S2m2Tensor* ins[4] = {f, g, z0, z1};
  if (int r = ew_check(out, ins, 4, "gate_blend")) return r;
  long long total4 = npix(*out) * (out->C / 4);
  launch_k(ew_kernel<EW_BLEND>, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*f), tv(*g), tv(*z0), tv(*z1), tv(*out), 0, total4);
  S2M2_LAUNCH_OK("gate_blend");
  return 0;
}

int s2m2b200_copy_channels(const S2m2Tensor* src, const S2m2Tensor* dst, void* stream) {
  S2M2_TRACE("copy_channels", src, dst);
  const S2m2Tensor* ins[1] = {src};
  if (int r = ew_check(dst, ins, 1, "copy_channels")) return r;
  long long total4 = npix(*dst) * (dst->C / 4);
  launch_k(ew_kernel<EW_COPY>, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*src), tv(*src), tv(*src), tv(*src), tv(*dst), 0, total4);
  S2M2_LAUNCH_OK("copy_channels");
  return 0;
}

int s2m2b200_unary(const S2m2Tensor* x, const S2m2Tensor* y, int32_t act, void* stream) {
  S2M2_TRACE("unary", x, y, act);
  const S2m2Tensor* ins[1] = {x};
  if (int r = ew_check(y, ins, 1, "unary")) return r;
  long long total4 = npix(*y) * (y->C / 4);
  launch_k(ew_kernel<EW_UNARY>, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*x), tv(*x), tv(*x), tv(*x), tv(*y), act, total4);
  S2M2_LAUNCH_OK("unary");
  return 0;
}

int s2m2b200_mul(const S2m2Tensor* a, const S2m2Tensor* b, const S2m2Tensor* out, void* stream) {
  S2M2_TRACE("mul", a, b, out);
  const S2m2Tensor* ins[2] = {a, b};
  if (int r = ew_check(out, ins, 2, "mul")) return r;
  long long total4 = npix(*out) * (out->C / 4);
  launch_k(ew_kernel<EW_MUL>, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*a), tv(*b), tv(*a), tv(*a), tv(*out), 0, total4);
  S2M2_LAUNCH_OK("mul");
  return 0;
}

int s2m2b200_gru_update(const S2m2Tensor* z, const S2m2Tensor* q, const S2m2Tensor* h, void* stream) {
  S2M2_TRACE("gru_update", z, q, h);
  const S2m2Tensor* ins[2] = {z, q};
  if (int r = ew_check(h, ins, 2, "gru_update")) return r;
  long long total4 = npix(*h) * (h->C / 4);
  launch_k(ew_kernel<EW_GRU>, grid_for(total4), 256, 0, (cudaStream_t)stream, tv(*z), tv(*q), tv(*z), tv(*z), tv(*h), 0, total4);
  S2M2_LAUNCH_OK("gru_update");
  return 0;
}

int s2m2b200_preprocess(const float* left, const float* right, int32_t B, int32_t H, int32_t W, const S2m2Tensor* out,
                        void* stream) {
  S2M2_TRACE("preprocess", left, right, B, H, W, out);
  return preprocess_any(left, right, B, H, W, out, stream);
}
int s2m2b200_preprocess_u8(const uint8_t* left, const uint8_t* right, int32_t B, int32_t H, int32_t W, const S2m2Tensor* out,
                           void* stream) {
  S2M2_TRACE("preprocess_u8", left, right, B, H, W, out);
  return preprocess_any(left, right, B, H, W, out, stream);
}

int s2m2b200_stem(const float* left, const float* right, int32_t B, int32_t H, int32_t W, const float* w0, const float* b0,
                  const float* w1, const float* b1, const S2m2Tensor* out, void* stream) {
  S2M2_TRACE("stem", left, right, B, H, W, w0, b0, w1, b1, out);
  return stem_any(left, right, B, H, W, w0, b0, w1, b1, out, stream);
}
int s2m2b200_stem_u8(const uint8_t* left, const uint8_t* right, int32_t B, int32_t H, int32_t W, const float* w0, const float* b0,
                     const float* w1, const float* b1, const S2m2Tensor* out, void* stream) {
  S2M2_TRACE("stem_u8", left, right, B, H, W, w0, b0, w1, b1, out);
  return stem_any(left, right, B, H, W, w0, b0, w1, b1, out, stream);
}

int s2m2b200_mask1x_inputs(const float* up, const S2m2Tensor* rgb, const float* wd, const float* bd, const float* wr,
                           const float* br, const S2m2Tensor* out, void* stream) {
  S2M2_TRACE("mask1x_inputs", up, rgb, wd, bd, wr, br, out);
  S2M2_CHECK(rgb->C == 3 && out->C == 32 && npix(*rgb) == npix(*out) && vec4_ok(*out) && aligned(up, 16),
             "mask1x_inputs: rgb (B,H,W,3), out a 32-channel 4-aligned view, up 16-byte aligned");
  Mask1xW hw;   // wd/bd/wr/br are HOST pointers (2.4 KB of weights as a kernel parameter)
  memcpy(hw.wd, wd, sizeof(hw.wd)); memcpy(hw.bd, bd, sizeof(hw.bd)); memcpy(hw.wr, wr, sizeof(hw.wr)); memcpy(hw.br, br, sizeof(hw.br));
  launch_k(mask1x_inputs_kernel, grid_for(npix(*out)), 256, 0, (cudaStream_t)stream, (const float4*)up, tv(*rgb), rgb->N, rgb->H, rgb->W,
                                                                                  hw, tv(*out));
  S2M2_LAUNCH_OK("mask1x_inputs");
  return 0;
}

int s2m2b200_global_refiner_prep(const float* disp, const float* conf, int64_t n, const S2m2Tensor* out, void* stream) {
  S2M2_TRACE("global_refiner_prep", disp, conf, n, out);
  S2M2_CHECK(out->C == 2 && npix(*out) == n, "global_refiner_prep: output must be a 2-channel view of npix pixels");
  launch_k(gr_prep_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, disp, conf, n, tv(*out));
  S2M2_LAUNCH_OK("global_refiner_prep");
  return 0;
}

int s2m2b200_global_refiner_merge(float* disp, const float* conf, const float* upd, int64_t upd_ld, int64_t n, int32_t pos,
                                  void* stream) {
  S2M2_TRACE("global_refiner_merge", disp, conf, upd, upd_ld, n, pos);
  S2M2_CHECK(upd_ld >= 1, "global_refiner_merge: bad upd_ld");
  launch_k(gr_merge_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, disp, conf, upd, upd_ld, n, pos);
  S2M2_LAUNCH_OK("global_refiner_merge");
  return 0;
}

int s2m2b200_local_refiner_prep(const float* disp, const float* conf, const float* occ, int64_t n, float* small, void* stream) {
  S2M2_TRACE("local_refiner_prep", disp, conf, occ, n, small);
  S2M2_CHECK(aligned(small, 16), "local_refiner_prep: small must be 16-byte aligned");
  launch_k(lr_prep_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, disp, conf, occ, n, (float4*)small);
  S2M2_LAUNCH_OK("local_refiner_prep");
  return 0;
}

int s2m2b200_local_refiner_update(float* disp, float* conf, float* occ, const float* small, const float* upd, int64_t upd_ld,
                                  int32_t B, int32_t h, int32_t w, int32_t pos, void* stream) {
  S2M2_TRACE("local_refiner_update", disp, conf, occ, small, upd, upd_ld, B, h, w, pos);
  long long n = (long long)B * h * w;
  S2M2_CHECK(upd_ld >= 3, "local_refiner_update: upd needs >= 3 channels");
  launch_k(lr_update_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, disp, conf, occ, (const float4*)small, upd, upd_ld, n, w, pos);
  S2M2_LAUNCH_OK("local_refiner_update");
  return 0;
}

}  // extern "C"
