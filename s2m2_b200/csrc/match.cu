// Matching-stage kernels on the fp32 scan-line cost volume: dust-bin log-Sinkhorn + argmax + soft-argmax,
// cost-volume lookup, convex upsampling.  HBM/L2-bound work; warp-shuffle reductions, no tensor cores.
// Reference: submodules.py:7-60,147-243; s2m2.py:101-133; utils.py:9-20.
#include <stdlib.h>

#include "common.cuh"

namespace s2m2 {

// ------------------------------------------------------------------------------------------------
// Sinkhorn + argmax + 5-tap soft-argmax.  One CTA per scan-line (b, y): the (w x w) row matrix is
// streamed from L2 once per half-iteration (2*iters + 1 passes); u, v live in shared memory.
//   A(i,j) = cv[i,j] (or -1e4 where j > i when use_positivity), dust-bin row/column = 0.
// ------------------------------------------------------------------------------------------------
constexpr int SK_THREADS = 512;   // 2 CTAs/SM x 148 SMs >= the 256 scan-lines of L@1216x1024: one wave
constexpr int SK_WARPS = SK_THREADS / 32;

__device__ __forceinline__ float sk_a(const float* __restrict__ row_i, int i, int j, int w, int pos) {
  // caller guarantees i < w
  if (j >= w) return 0.f;
  if (pos && j > i) return -1e4f;
  return row_i[j];
}

__global__ void __launch_bounds__(SK_THREADS) sinkhorn_kernel(const float* __restrict__ cv, int w, int pos, int iters,
                                                               float* __restrict__ disp, float* __restrict__ conf,
                                                               float* __restrict__ occ, int* __restrict__ amax) {
  pdl_sync();
  extern __shared__ float sh[];
  const int W1 = w + 1;
  float* u = sh;                 // [W1]
  float* v = sh + W1;            // [W1]
  float* red_m = v + W1;         // [SK_WARPS][32]
  float* red_s = red_m + SK_WARPS * 32;
  const long long line = blockIdx.x;
  const float* A = cv + line * (long long)w * w;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float log_m_in = logf(1.0f / (2.0f * (float)w));            // log(1/(2w))   (submodules.py:187-189)
  const float log_m_dust = logf((float)w / (2.0f * (float)w));       // log(w/(2w))
  const float log_const = logf(2.0f * (float)w);                      // submodules.py:198-199

  for (int i = threadIdx.x; i < W1; i += SK_THREADS) { u[i] = 0.f; v[i] = 0.f; }
  __syncthreads();

  // column pass: v[j] = log_nu[j] - LSE_i (A(i,j) + u[i])
  auto col_pass = [&](bool use_u) {
    for (int jc = 0; jc < W1; jc += 32) {
      const int j = jc + lane;
      float m = -1e30f, s = 0.f;
      if (j < W1) {
        for (int i = wid; i < W1; i += SK_WARPS) {
          float a = (i < w) ? sk_a(A + (long long)i * w, i, j, w, pos) : 0.f;
          if (use_u) a += u[i];
          if (a > m) { s = s * expf(m - a) + 1.f; m = a; }
          else s += expf(a - m);
        }
      }
      red_m[wid * 32 + lane] = m;
      red_s[wid * 32 + lane] = s;
      __syncthreads();
      if (wid == 0 && j < W1) {
        float M = red_m[lane];
#pragma unroll
        for (int k = 1; k < SK_WARPS; ++k) M = fmaxf(M, red_m[k * 32 + lane]);
        float S = 0.f;
#pragma unroll
        for (int k = 0; k < SK_WARPS; ++k) S += red_s[k * 32 + lane] * expf(red_m[k * 32 + lane] - M);
        float lse = M + logf(fmaxf(S, 1e-30f));
        v[j] = (j < w ? log_m_in : log_m_dust) - lse;
      }
      __syncthreads();
    }
  };
  // row pass: u[i] = log_mu[i] - LSE_j (A(i,j) + v[j])
  auto row_pass = [&]() {
    for (int i = wid; i < W1; i += SK_WARPS) {
      const float* row = A + (long long)i * w;
      float m = -1e30f;
      for (int j = lane; j < W1; j += 32) {
        float a = ((i < w) ? sk_a(row, i, j, w, pos) : 0.f) + v[j];
        m = fmaxf(m, a);
      }
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < W1; j += 32) {
        float a = ((i < w) ? sk_a(row, i, j, w, pos) : 0.f) + v[j];
        s += expf(a - m);
      }
      s = warp_sum(s);
      if (lane == 0) u[i] = (i < w ? log_m_in : log_m_dust) - (m + logf(fmaxf(s, 1e-30f)));
    }
    __syncthreads();
  };

  col_pass(false);
  row_pass();
  for (int it = 1; it < iters; ++it) {
    col_pass(true);
    row_pass();
  }

  // final: prob, row sum (occ), argmax, 5-tap window
  for (int i = wid; i < w; i += SK_WARPS) {
    const float* row = A + (long long)i * w;
    const float ui = u[i] + log_const;
    float best = -1.f, sum = 0.f;
    int bidx = 0x7fffffff;
    const int jend = pos ? i + 1 : w;  // masked entries are exactly 0 (submodules.py:223)
    for (int j = lane; j < w; j += 32) {
      float pj = (j < jend) ? expf(row[j] + ui + v[j]) : 0.f;
      sum += pj;
      if (pj > best) { best = pj; bidx = j; }
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    // 5 taps around the arg-max, zero outside [0, w)  (submodules.py:228-236)
    float tap = 0.f;
    {
      int j = bidx - 2 + lane;
      if (lane < 5 && j >= 0 && j < jend) tap = expf(row[j] + ui + v[j]);
    }
    float cf = 0.f, cr = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float wk = __shfl_sync(0xffffffffu, tap, k);
      cf += wk;
      cr += wk * (float)(bidx + k - 2);
    }
    if (lane == 0) {
      cr = (cr + 1e-4f) / (cf + 1e-4f);
      long long o = line * w + i;
      disp[o] = (float)i - cr;
      conf[o] = cf;
      occ[o] = sum;
      amax[o] = bidx;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cost-volume lookup: 9 taps on the row cv[b,y,x,:] at x-disp+dx and 9 on the width-pooled row
// ------------------------------------------------------------------------------------------------
__global__ void cv_lookup_kernel(const float* __restrict__ cv, const float* __restrict__ disp, long long npix, int w,
                                 float* __restrict__ out) {
  pdl_sync();
  const long long total = npix * 18;
  const int w2 = w / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i / 18;
    int t = (int)(i - pix * 18);
    int x = (int)(pix % w);
    const float* row = cv + pix * w;
    float d = disp[pix];
    float r;
    if (t < 9) {
      float pos = (float)x - d + (float)(t - 4);
      float p0 = floorf(pos);
      float f = pos - p0;
      int i0 = (int)p0, i1 = i0 + 1;
      float v0 = (i0 >= 0 && i0 < w) ? row[i0] : 0.f;
      float v1 = (i1 >= 0 && i1 < w) ? row[i1] : 0.f;
      r = (1.f - f) * v0 + f * v1;
    } else {
      float pos = (float)x / 2.f - d / 2.f + (float)(t - 13);
      float p0 = floorf(pos);
      float f = pos - p0;
      int i0 = (int)p0, i1 = i0 + 1;
      float v0 = (i0 >= 0 && i0 < w2) ? (row[2 * i0] + row[2 * i0 + 1]) * 0.5f : 0.f;
      float v1 = (i1 >= 0 && i1 < w2) ? (row[2 * i1] + row[2 * i1 + 1]) * 0.5f : 0.f;
      r = (1.f - f) * v0 + f * v1;
    }
    out[i] = r / 16.f;
  }
}

// ------------------------------------------------------------------------------------------------
// convex upsampling
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax9(const float* __restrict__ l, float* wt) {
  float m = l[0];
#pragma unroll
  for (int k = 1; k < 9; ++k) m = fmaxf(m, l[k]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { wt[k] = expf(l[k] - m); s += wt[k]; }
  float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < 9; ++k) wt[k] *= inv;
}

__global__ void upsample4x_kernel(const float* __restrict__ disp, const float* __restrict__ occ, const float* __restrict__ conf,
                                  const float* __restrict__ logits, int lld, int B, int h, int w, float4* __restrict__ up) {
  pdl_sync();
  const int H = 4 * h, W = 4 * w;
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % W);
    long long t = i / W;
    int Y = (int)(t % H);
    long long b = t / H;
    float wt[9];
    softmax9(logits + i * lld, wt);
    int cy = Y >> 2, cx = X >> 2;
    float d = 0.f, o = 0.f, c = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int yy = min(max(cy + dy, 0), h - 1), xx = min(max(cx + dx, 0), w - 1);
        long long q = (b * h + yy) * w + xx;
        float wk = wt[(dy + 1) * 3 + dx + 1];
        d += (disp[q] * 4.f) * wk;
        o += occ[q] * wk;
        c += conf[q] * wk;
      }
    up[i] = make_float4(d, o, c, 0.f);
  }
}

// (y0, x0, Hc, Wc): the output window — image_crop (image_utils.py:66-96) fused into the last kernel of the forward: only the
// un-padded pixels are filtered and written, as three dense (B, Hc, Wc) maps.  The full frame is y0 = x0 = 0, Hc = H, Wc = W.
__global__ void filter1x_kernel(const float4* __restrict__ up, const float* __restrict__ logits, int lld, int B, int H, int W,
                                float* __restrict__ out, int y0, int x0, int Hc, int Wc) {
  pdl_sync();
  const long long total = (long long)B * Hc * Wc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % Wc) + x0;
    long long t = i / Wc;
    int Y = (int)(t % Hc) + y0;
    long long b = t / Hc;
    float wt[9];
    softmax9(logits + ((b * H + Y) * W + X) * lld, wt);
    float d = 0.f, o = 0.f, c = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int yy = min(max(Y + dy, 0), H - 1), xx = min(max(X + dx, 0), W - 1);
        float4 v = up[(b * H + yy) * W + xx];
        float wk = wt[(dy + 1) * 3 + dx + 1];
        d += v.x * wk;
        o += v.y * wk;
        c += v.z * wk;
      }
    out[i] = d;
    out[total + i] = o;
    out[2 * total + i] = c;
  }
}

// output_upsample=True branch (s2m2.py:123-127,194-195): the 3x3 neighbourhood is taken at (Y/2, X/2) (nearest x2), the
// 9 logits are bilinearly interpolated x2 (align_corners=False) BEFORE the softmax, disparity is doubled.
__global__ void filter2x_kernel(const float4* __restrict__ up, const float* __restrict__ logits, int lld, int B, int H, int W,
                                float* __restrict__ out) {
  pdl_sync();
  const int H2 = 2 * H, W2 = 2 * W;
  const long long total = (long long)B * H2 * W2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % W2);
    long long t = i / W2;
    int Y = (int)(t % H2);
    long long b = t / H2;
    const float sy = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* l00 = logits + ((b * H + y0) * W + x0) * lld;
    const float* l01 = logits + ((b * H + y0) * W + x1) * lld;
    const float* l10 = logits + ((b * H + y1) * W + x0) * lld;
    const float* l11 = logits + ((b * H + y1) * W + x1) * lld;
    float lg[9], wt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) lg[k] = hy * (hx * l00[k] + lx * l01[k]) + ly * (hx * l10[k] + lx * l11[k]);
    softmax9(lg, wt);
    const int cy = Y >> 1, cx = X >> 1;
    float d = 0.f, o = 0.f, c = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int yy = min(max(cy + dy, 0), H - 1), xx = min(max(cx + dx, 0), W - 1);
        float4 v = up[(b * H + yy) * W + xx];
        float wk = wt[(dy + 1) * 3 + dx + 1];
        d += v.x * wk;
        o += v.y * wk;
        c += v.z * wk;
      }
    out[i] = 2.f * d;
    out[total + i] = o;
    out[2 * total + i] = c;
  }
}


// ------------------------------------------------------------------------------------------------
// consumers of the three output maps, fused so that the maps need not leave the device
// ------------------------------------------------------------------------------------------------
// mean confidence over the interior of the centre crop: image_crop (image_utils.py:66-96) to (crop_h, crop_w), then
// [margin:-margin, margin:-margin].mean() (model_utils.py:93-94).  One CTA per image, fixed summation order (deterministic).
__global__ void __launch_bounds__(1024) confidence_score_kernel(const float* __restrict__ conf, int H, int W, int y0, int y1, int x0, int x1,
                                                                float* __restrict__ out) {
  pdl_sync();
  __shared__ double sh[1024];
  const float* img = conf + (long long)blockIdx.x * H * W;
  const int ww = x1 - x0, n = (y1 - y0) * ww;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = y0 + i / ww, x = x0 + i - (i / ww) * ww;
    acc += (double)img[(long long)y * W + x];
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = n > 0 ? (float)(sh[0] / n) : nanf("");
}

// disparity -> camera-frame points (model_utils.py:111-136 + open3d's RGBD back-projection): depth = baseline*fx/(disp+doffs),
// depth[disp<=0] = 1e9, z = depth/depth_scale, dropped if z > depth_trunc; x = (u-cx) z / fx, y = (v-cy) z / fx;
// trust mask conf > .1 and occ > .5 (vis_utils.py:62).  xyz: (B,H,W,3) fp32, valid: (B,H,W) bytes.
__global__ void disparity_to_points_kernel(const float* __restrict__ disp, const float* __restrict__ conf, const float* __restrict__ occ,
                                           long long n, int H, int W, float fx, float cx, float cy, float baseline, float doffs,
                                           float depth_scale, float depth_trunc, float* __restrict__ xyz, uint8_t* __restrict__ valid) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % W), v = (int)((i / W) % H);
    const float d = disp[i];
    float depth = baseline * fx / (d + doffs);
    if (d <= 0.f) depth = 1e9f;
    const float z = depth / depth_scale;
    bool ok = z > 0.f && z <= depth_trunc;
    if (conf) ok = ok && conf[i] > 0.1f;
    if (occ) ok = ok && occ[i] > 0.5f;
    const float zz = (z > 0.f && z <= depth_trunc) ? z : 0.f;
    xyz[3 * i + 0] = (u - cx) * zz / fx;
    xyz[3 * i + 1] = (v - cy) * zz / fx;
    xyz[3 * i + 2] = zz;
    valid[i] = ok ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// image_pad (image_utils.py:27-63), the pre-processing of run_stereo_matching (model_utils.py:72-73), on the device:
// centre zero-padding to multiples of `factor`, the padding then filled with a blurred copy of the padded image
// (adaptive_avg_pool2d to (H/factor, W/factor) -> bilinear back up, align_corners=False), the original pasted over it.
//   kernel 1  pad_pool_kernel   one warp per (image, channel, pooled cell): mean of its adaptive window of the zero-padded image
//   kernel 2  pad_fill_kernel   every output pixel: inside -> the input pixel, padding -> bilinear sample of the pooled grid
// Input uint8 or fp32 (the same integer values either way), output fp32 (B,3,Hn,Wn).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pad_pool_kernel(const T* __restrict__ img, int BC, int H, int W, int Hn, int Wn, int hs, int ws, int Ho, int Wo,
                                float* __restrict__ pooled) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long cells = (long long)BC * Ho * Wo;
  for (long long c = warp; c < cells; c += nw) {
    const int ox = (int)(c % Wo);
    const long long t = c / Wo;
    const int oy = (int)(t % Ho);
    const long long bc = t / Ho;
    // adaptive windows of the PADDED image: [floor(o*In/Out), ceil((o+1)*In/Out))
    const int y0 = (int)(((long long)oy * Hn) / Ho), y1 = (int)((((long long)oy + 1) * Hn + Ho - 1) / Ho);
    const int x0 = (int)(((long long)ox * Wn) / Wo), x1 = (int)((((long long)ox + 1) * Wn + Wo - 1) / Wo);
    const int ww = x1 - x0, n = (y1 - y0) * ww;
    float s = 0.f;
    for (int k = lane; k < n; k += 32) {
      const int yy = y0 + k / ww - hs, xx = x0 + k % ww - ws;          // coordinates in the un-padded image
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += (float)img[(bc * H + yy) * W + xx];
    }
    s = warp_sum(s);
    if (lane == 0) pooled[c] = s / (float)n;
  }
}

template <typename T>
__global__ void pad_fill_kernel(const T* __restrict__ img, const float* __restrict__ pooled, int BC, int H, int W, int Hn, int Wn, int hs,
                                int ws, int Ho, int Wo, float* __restrict__ out) {
  pdl_sync();
  const long long total = (long long)BC * Hn * Wn;
  const float sy = (float)Ho / (float)Hn, sx = (float)Wo / (float)Wn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % Wn);
    const long long t = i / Wn;
    const int Y = (int)(t % Hn);
    const long long bc = t / Hn;
    const int yy = Y - hs, xx = X - ws;
    float v;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      v = (float)img[(bc * H + yy) * W + xx];
    } else {   // F.interpolate(bilinear, align_corners=False): src = (dst + .5) * scale - .5, clamped at 0
      const float fy = fmaxf(((float)Y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf(((float)X + 0.5f) * sx - 0.5f, 0.f);
      const int iy = (int)fy, ix = (int)fx;
      const int iy1 = min(iy + 1, Ho - 1), ix1 = min(ix + 1, Wo - 1);
      const float ty = fy - (float)iy, tx = fx - (float)ix;
      const float* g = pooled + bc * Ho * Wo;
      const float a = g[iy * Wo + ix], b2 = g[iy * Wo + ix1], c2 = g[iy1 * Wo + ix], d2 = g[iy1 * Wo + ix1];
      v = (1.f - ty) * ((1.f - tx) * a + tx * b2) + ty * ((1.f - tx) * c2 + tx * d2);
    }
    out[i] = v;
  }
}

// ---- stereo rectification: cv2.remap(uint8, INTER_LINEAR, BORDER_CONSTANT 0), bit-exact ------------------------------------
// image_utils.py:108-136 (rectify_images) applies cv2.remap with float32 maps to both uint8 images; the calibration optimisers do
// that for every candidate rotation (base.py:15-36, cem.py:64-72).  OpenCV's 8-bit bilinear remap is fixed point: the map is
// rounded to 1/32 pixel (round-half-even of x*32), the four weights are (32-fy)(32-fx)*32 ... on a 2^15 scale, the result is
// (sum + 2^14) >> 15; taps outside the source read the border value 0.  At integer coordinates the table entry is
// [32767, 0, 0, 1] (saturate_cast<short>(32768) plus OpenCV's sum correction).  dst is planar (N,3,Hd,Wd) — what the forward's
// uint8 ingest takes — or interleaved (N,Hd,Wd,3) like cv2's output.
__global__ void remap_u8_kernel(const uint8_t* __restrict__ src, long long src_stride, int Hs, int Ws, const float* __restrict__ mapx,
                                const float* __restrict__ mapy, long long n_pix, int Hd, int Wd, uint8_t* __restrict__ dst, int planar) {
  pdl_sync();
  const long long hw = (long long)Hd * Wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, r = i - n * hw;
    const int sx = __float2int_rn(mapx[i] * 32.f), sy = __float2int_rn(mapy[i] * 32.f);
    int ix = sx >> 5, iy = sy >> 5;
    ix = ix < -32768 ? -32768 : (ix > 32767 ? 32767 : ix);     // saturate_cast<short>
    iy = iy < -32768 ? -32768 : (iy > 32767 ? 32767 : iy);
    const int fx = sx & 31, fy = sy & 31;
    int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }
    const uint8_t* S = src + n * src_stride;
    int acc[3] = {0, 0, 0};
    const bool y0 = iy >= 0 && iy < Hs, y1 = iy + 1 >= 0 && iy + 1 < Hs, x0 = ix >= 0 && ix < Ws, x1 = ix + 1 >= 0 && ix + 1 < Ws;
    if (y0 && x0) { const uint8_t* p = S + ((long long)iy * Ws + ix) * 3; acc[0] += p[0] * w00; acc[1] += p[1] * w00; acc[2] += p[2] * w00; }
    if (y0 && x1) { const uint8_t* p = S + ((long long)iy * Ws + ix + 1) * 3; acc[0] += p[0] * w01; acc[1] += p[1] * w01; acc[2] += p[2] * w01; }
    if (y1 && x0) { const uint8_t* p = S + ((long long)(iy + 1) * Ws + ix) * 3; acc[0] += p[0] * w10; acc[1] += p[1] * w10; acc[2] += p[2] * w10; }
    if (y1 && x1) { const uint8_t* p = S + ((long long)(iy + 1) * Ws + ix + 1) * 3; acc[0] += p[0] * w11; acc[1] += p[1] * w11; acc[2] += p[2] * w11; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int v = (acc[c] + (1 << 14)) >> 15;
      v = v > 255 ? 255 : v;
      if (planar) dst[(n * 3 + c) * hw + r] = (uint8_t)v;
      else dst[i * 3 + c] = (uint8_t)v;
    }
  }
}

}  // namespace s2m2

namespace s2m2 {
int sinkhorn_cluster_launch(const float* cv, int B, int h, int w, float* disp, float* conf, float* occ, int* amax, int pos, int iters,
                            cudaStream_t s);
}
using namespace s2m2;

extern "C" {

int s2m2b200_sinkhorn_softargmax(const float* cv, int32_t B, int32_t h, int32_t w, float* disp, float* conf, float* occ,
                                 int32_t* amax, int32_t use_positivity, int32_t iters, void* stream) {
  S2M2_TRACE("sinkhorn_softargmax", cv, B, h, w, disp, conf, occ, amax, use_positivity, iters);
  S2M2_CHECK(w >= 8 && w <= 5000 && iters >= 1, "sinkhorn: bad w=%d / iters=%d (8 <= w <= 5000)", w, iters);
  // the matrix-resident cluster kernel (sinkhorn_cluster.cu) takes every width whose scan-line fits a cluster's shared memory
  // (w <= 639 with positivity); S2M2_SK_IMPL=1 forces the streaming kernel below
  static const int impl = getenv("S2M2_SK_IMPL") ? atoi(getenv("S2M2_SK_IMPL")) : 0;
  if (impl != 1) {
    const int r = sinkhorn_cluster_launch(cv, B, h, w, disp, conf, occ, amax, use_positivity, iters, (cudaStream_t)stream);
    if (r != 1) return r;
  }
  size_t smem = sizeof(float) * (2 * (w + 1) + 2 * SK_WARPS * 32);   // <= 44 KB for w <= 5000: below the 48 KB default limit
  launch_k(sinkhorn_kernel, B * h, SK_THREADS, smem, (cudaStream_t)stream, cv, w, use_positivity, iters, disp, conf, occ, amax);
  S2M2_LAUNCH_OK("sinkhorn_softargmax");
  return 0;
}

int s2m2b200_cv_lookup(const float* cv, const float* disp, int32_t B, int32_t h, int32_t w, float* out, void* stream) {
  S2M2_TRACE("cv_lookup", cv, disp, B, h, w, out);
  long long n = (long long)B * h * w;
  long long g = (n * 18 + 255) / 256;
  launch_k(cv_lookup_kernel, (int)(g > 148 * 64 ? 148 * 64 : g), 256, 0, (cudaStream_t)stream, cv, disp, n, w, out);
  S2M2_LAUNCH_OK("cv_lookup");
  return 0;
}

int s2m2b200_upsample4x_convex(const float* disp, const float* occ, const float* conf, const float* logits, int32_t logits_ld,
                               int32_t B, int32_t h, int32_t w, float* up, void* stream) {
  S2M2_TRACE("upsample4x_convex", disp, occ, conf, logits, logits_ld, B, h, w, up);
  S2M2_CHECK(aligned(up, 16) && logits_ld >= 9, "upsample4x_convex: up must be 16-byte aligned, logits_ld >= 9");
  long long n = 16LL * B * h * w;
  long long g = (n + 255) / 256;
  launch_k(upsample4x_kernel, (int)(g > 148 * 64 ? 148 * 64 : g), 256, 0, (cudaStream_t)stream, disp, occ, conf, logits, logits_ld, B, h, w, (float4*)up);
  S2M2_LAUNCH_OK("upsample4x_convex");
  return 0;
}

int s2m2b200_filter1x(const float* up, const float* logits, int32_t logits_ld, int32_t B, int32_t H, int32_t W, float* out,
                      int32_t upsample2, void* stream) {
  S2M2_TRACE("filter1x", up, logits, logits_ld, B, H, W, out, upsample2);
  S2M2_CHECK(aligned(up, 16) && logits_ld >= 9, "filter1x: up must be 16-byte aligned, logits_ld >= 9");
  if (upsample2) {
    long long n2 = 4LL * B * H * W;
    long long g2 = (n2 + 255) / 256;
    launch_k(filter2x_kernel, (int)(g2 > 148 * 64 ? 148 * 64 : g2), 256, 0, (cudaStream_t)stream, (const float4*)up, logits, logits_ld, B, H, W, out);
    S2M2_LAUNCH_OK("filter2x");
    return 0;
  }
  long long n = (long long)B * H * W;
  long long g = (n + 255) / 256;
  launch_k(filter1x_kernel, (int)(g > 148 * 64 ? 148 * 64 : g), 256, 0, (cudaStream_t)stream, (const float4*)up, logits, logits_ld, B, H, W, out,
           0, 0, H, W);
  S2M2_LAUNCH_OK("filter1x");
  return 0;
}

int s2m2b200_filter1x_crop(const float* up, const float* logits, int32_t logits_ld, int32_t B, int32_t H, int32_t W, int32_t crop_h,
                           int32_t crop_w, float* out, void* stream) {
  S2M2_TRACE("filter1x_crop", up, logits, logits_ld, B, H, W, crop_h, crop_w, out);
  S2M2_CHECK(aligned(up, 16) && logits_ld >= 9, "filter1x_crop: up must be 16-byte aligned, logits_ld >= 9");
  S2M2_CHECK(crop_h > 0 && crop_w > 0 && crop_h <= H && crop_w <= W, "filter1x_crop: bad crop %dx%d of %dx%d", crop_h, crop_w, H, W);
  const int y0 = (H - crop_h) / 2, x0 = (W - crop_w) / 2;          // image_crop drops floor(d/2) rows / columns in front
  long long n = (long long)B * crop_h * crop_w;
  long long g = (n + 255) / 256;
  launch_k(filter1x_kernel, (int)(g > 148 * 64 ? 148 * 64 : g), 256, 0, (cudaStream_t)stream, (const float4*)up, logits, logits_ld, B, H, W, out,
           y0, x0, crop_h, crop_w);
  S2M2_LAUNCH_OK("filter1x_crop");
  return 0;
}

int s2m2b200_image_pad(const void* img, int32_t is_u8, int32_t B, int32_t H, int32_t W, int32_t factor, float* out, float* work,
                       void* stream) {
  S2M2_TRACE("image_pad", img, is_u8, B, H, W, factor, out, work);
  S2M2_CHECK(factor > 0 && H >= factor && W >= factor, "image_pad: the image must be at least factor x factor");
  const int Hn = (H + factor - 1) / factor * factor, Wn = (W + factor - 1) / factor * factor;
  const int hs = (Hn - H) / 2, ws = (Wn - W) / 2;                   // image_utils.py:47-52: floor(pad/2) in front
  const int Ho = H / factor, Wo = W / factor;                        // image_utils.py:54 (the UN-padded size / factor)
  const int BC = B * 3;
  cudaStream_t s = (cudaStream_t)stream;
  const long long cells = (long long)BC * Ho * Wo, total = (long long)BC * Hn * Wn;
  const int gp = (int)((cells * 32 + 255) / 256 > 148 * 32 ? 148 * 32 : (cells * 32 + 255) / 256);
  const int gf = (int)((total + 255) / 256 > 148 * 64 ? 148 * 64 : (total + 255) / 256);
  if (is_u8) {
    launch_k(pad_pool_kernel<uint8_t>, gp, 256, 0, s, (const uint8_t*)img, BC, H, W, Hn, Wn, hs, ws, Ho, Wo, work);
    S2M2_LAUNCH_OK("pad_pool");
    launch_k(pad_fill_kernel<uint8_t>, gf, 256, 0, s, (const uint8_t*)img, (const float*)work, BC, H, W, Hn, Wn, hs, ws, Ho, Wo, out);
  } else {
    launch_k(pad_pool_kernel<float>, gp, 256, 0, s, (const float*)img, BC, H, W, Hn, Wn, hs, ws, Ho, Wo, work);
    S2M2_LAUNCH_OK("pad_pool");
    launch_k(pad_fill_kernel<float>, gf, 256, 0, s, (const float*)img, (const float*)work, BC, H, W, Hn, Wn, hs, ws, Ho, Wo, out);
  }
  S2M2_LAUNCH_OK("pad_fill");
  return 0;
}

}  // extern "C"

int s2m2b200_remap_u8(const uint8_t* src, int64_t src_stride, int32_t Hs, int32_t Ws, const float* mapx, const float* mapy, int32_t N,
                      int32_t Hd, int32_t Wd, uint8_t* dst, int32_t planar, void* stream) {
  S2M2_TRACE("remap_u8", src, src_stride, Hs, Ws, mapx, mapy, N, Hd, Wd, dst, planar);
  S2M2_CHECK(src && mapx && mapy && dst, "remap_u8: null pointer");
  S2M2_CHECK(Hs > 0 && Ws > 0 && N > 0 && Hd > 0 && Wd > 0 && src_stride >= 0, "remap_u8: bad geometry");
  const long long n = (long long)N * Hd * Wd, g = (n + 255) / 256;
  launch_k(remap_u8_kernel, (int)(g > 148 * 32 ? 148 * 32 : g), 256, 0, (cudaStream_t)stream, src, (long long)src_stride, Hs, Ws, mapx, mapy, n, Hd, Wd,
           dst, planar);
  S2M2_LAUNCH_OK("remap_u8");
  return 0;
}

int s2m2b200_confidence_score(const float* conf, int32_t B, int32_t H, int32_t W, int32_t crop_h, int32_t crop_w, int32_t margin,
                              float* out, void* stream) {
  S2M2_TRACE("confidence_score", conf, B, H, W, crop_h, crop_w, margin, out);
  S2M2_CHECK(crop_h > 0 && crop_w > 0 && crop_h <= H && crop_w <= W && margin >= 0, "confidence_score: bad crop / margin");
  const int ys = (H - crop_h) / 2, xs = (W - crop_w) / 2;     // image_crop drops floor(d/2) rows/columns in front
  const int y0 = ys + margin, y1 = ys + crop_h - margin, x0 = xs + margin, x1 = xs + crop_w - margin;
  S2M2_CHECK(y1 > y0 && x1 > x0, "confidence_score: the margin leaves no interior");
  launch_k(confidence_score_kernel, B, 1024, 0, (cudaStream_t)stream, conf, H, W, y0, y1, x0, x1, out);
  S2M2_LAUNCH_OK("confidence_score");
  return 0;
}

int s2m2b200_disparity_to_points(const float* disp, const float* conf, const float* occ, int32_t B, int32_t H, int32_t W, float fx, float cx,
                                 float cy, float baseline, float doffs, float depth_scale, float depth_trunc, float* xyz, uint8_t* valid,
                                 void* stream) {
  S2M2_TRACE("disparity_to_points", disp, conf, occ, B, H, W, fx, cx, cy, baseline, doffs, depth_scale, depth_trunc, xyz, valid);
  S2M2_CHECK(fx > 0.f && depth_scale > 0.f, "disparity_to_points: fx and depth_scale must be positive");
  const long long n = (long long)B * H * W;
  const long long g = (n + 255) / 256;
  launch_k(disparity_to_points_kernel, (int)(g > 148 * 32 ? 148 * 32 : g), 256, 0, (cudaStream_t)stream, disp, conf, occ, n, H, W, fx, cx, cy,
           baseline, doffs, depth_scale, depth_trunc, xyz, valid);
  S2M2_LAUNCH_OK("disparity_to_points");
  return 0;
}
