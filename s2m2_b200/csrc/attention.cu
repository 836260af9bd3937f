// Multi-head attention, CUDA-core fp32 flash-style kernel (parity path; also the fallback for head dims the
// tensor-core kernel does not take).  One CTA = 32 queries of one (sequence, head); K/V streamed through shared
// memory in tiles of 32 keys with an online softmax; optional contextual relative-PE accumulation
// (attentions.py:42-48) from two tiny tables instead of the (N,N,32) tensor of utils.py:32-60.
#include <stdlib.h>

#include "attention.cuh"

namespace s2m2 {

constexpr int AT_BQ = 32, AT_BK = 32, AT_THREADS = 256, AT_QPW = 4;  // 8 warps x 4 queries

template <int MAXI>  // MAXI = ceil(d / 32)
__global__ void __launch_bounds__(AT_THREADS) attention_simt_kernel(AttnP p) {
  pdl_sync();
  extern __shared__ __align__(16) float smem[];
  const int d = p.d, dp = d + 4;  // padded row (keeps 16 B alignment, conflict-free float4 rows)
  float* Qs = smem;               // [AT_BQ][dp]
  float* Ks = Qs + AT_BQ * dp;    // [AT_BK][dp]
  float* Vs = Ks + AT_BK * dp;    // [AT_BK][d]
  const int q0 = blockIdx.x * AT_BQ, head = blockIdx.y, s = blockIdx.z;
  const int s_kv = (s + p.kv_shift) % p.S;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool vec = (d % 4 == 0);

  // load the query tile (zero rows beyond N)
  for (int e = tid; e < AT_BQ * (vec ? d / 4 : d); e += AT_THREADS) {
    if (vec) {
      int r = e / (d / 4), c = (e % (d / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + r < p.N) v = ld4(p.qkv, ((long long)s * p.N + q0 + r) * p.qkv_ld + head * d + c, p.qh);
      *(float4*)&Qs[r * dp + c] = v;
    } else {
      int r = e / d, c = e % d;
      Qs[r * dp + c] = (q0 + r < p.N) ? ld1(p.qkv, ((long long)s * p.N + q0 + r) * p.qkv_ld + head * d + c, p.qh) : 0.f;
    }
  }

  float acc[AT_QPW][MAXI];
  float pacc[AT_QPW];
  float mrow[AT_QPW], lrow[AT_QPW];
#pragma unroll
  for (int q = 0; q < AT_QPW; ++q) {
    mrow[q] = -1e30f; lrow[q] = 0.f; pacc[q] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) acc[q][i] = 0.f;
  }
  // token coordinates of this warp's queries (for the relative PE)
  int qx[AT_QPW], qy[AT_QPW];
#pragma unroll
  for (int q = 0; q < AT_QPW; ++q) {
    int t = q0 + wid * AT_QPW + q;
    if (t >= p.N) t = 0;  // padded query rows: keep the table index in range
    qx[q] = p.pe_w ? t % p.pe_w : 0;
    qy[q] = p.pe_w ? t / p.pe_w : 0;
  }

  for (int k0 = 0; k0 < p.N; k0 += AT_BK) {
    __syncthreads();
    for (int e = tid; e < AT_BK * (vec ? d / 4 : d); e += AT_THREADS) {
      if (vec) {
        int r = e / (d / 4), c = (e % (d / 4)) * 4;
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (k0 + r < p.N) {
          long long base = ((long long)s_kv * p.N + k0 + r) * p.qkv_ld + head * d + c;
          kv = ld4(p.qkv, base + p.C, p.qh);
          vv = ld4(p.qkv, base + 2 * p.C, p.qh);
        }
        *(float4*)&Ks[r * dp + c] = kv;
        *(float4*)&Vs[r * d + c] = vv;
      } else {
        int r = e / d, c = e % d;
        float kv = 0.f, vv = 0.f;
        if (k0 + r < p.N) {
          long long base = ((long long)s_kv * p.N + k0 + r) * p.qkv_ld + head * d + c;
          kv = ld1(p.qkv, base + p.C, p.qh);
          vv = ld1(p.qkv, base + 2 * p.C, p.qh);
        }
        Ks[r * dp + c] = kv;
        Vs[r * d + c] = vv;
      }
    }
    __syncthreads();

    // scores: lane = key, 4 queries per warp
    float sc[AT_QPW];
#pragma unroll
    for (int q = 0; q < AT_QPW; ++q) sc[q] = 0.f;
    const float* krow = Ks + lane * dp;
    if (vec) {
      for (int c = 0; c < d; c += 4) {
        float4 kk = *(const float4*)&krow[c];
#pragma unroll
        for (int q = 0; q < AT_QPW; ++q) {
          float4 qq = *(const float4*)&Qs[(wid * AT_QPW + q) * dp + c];
          sc[q] = fmaf(qq.x, kk.x, sc[q]); sc[q] = fmaf(qq.y, kk.y, sc[q]);
          sc[q] = fmaf(qq.z, kk.z, sc[q]); sc[q] = fmaf(qq.w, kk.w, sc[q]);
        }
      }
    } else {
      for (int c = 0; c < d; ++c) {
        float kk = krow[c];
#pragma unroll
        for (int q = 0; q < AT_QPW; ++q) sc[q] = fmaf(Qs[(wid * AT_QPW + q) * dp + c], kk, sc[q]);
      }
    }
    const bool kvalid = (k0 + lane) < p.N;
    float pr[AT_QPW];
#pragma unroll
    for (int q = 0; q < AT_QPW; ++q) {
      float sv = kvalid ? sc[q] * p.scale : -1e30f;
      float mnew = fmaxf(mrow[q], warp_max(sv));
      float corr = expf(mrow[q] - mnew);
      float pv = kvalid ? expf(sv - mnew) : 0.f;
      lrow[q] = lrow[q] * corr + warp_sum(pv);
      mrow[q] = mnew;
      pr[q] = pv;
      pacc[q] *= corr;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) acc[q][i] *= corr;
    }
    // O += P V   (lane owns channels lane + 32 i); PE: lane owns one of the 32 PE channels
    const int kmax = min(AT_BK, p.N - k0);
    for (int j = 0; j < kmax; ++j) {
      float pj[AT_QPW];
#pragma unroll
      for (int q = 0; q < AT_QPW; ++q) pj[q] = __shfl_sync(0xffffffffu, pr[q], j);
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        int c = lane + 32 * i;
        if (c < d) {
          float vv = Vs[j * d + c];
#pragma unroll
          for (int q = 0; q < AT_QPW; ++q) acc[q][i] = fmaf(pj[q], vv, acc[q][i]);
        }
      }
      if (p.pesum) {
        int t = k0 + j;
        int kx = t % p.pe_w, ky = t / p.pe_w;
#pragma unroll
        for (int q = 0; q < AT_QPW; ++q) {
          float tv = lane < 16 ? p.pe_x[(qx[q] - kx + p.pe_w - 1) * 16 + lane]
                               : p.pe_y[(qy[q] - ky + p.pe_h - 1) * 16 + lane - 16];
          pacc[q] = fmaf(pj[q], 0.5f * tv, pacc[q]);
        }
      }
    }
  }

#pragma unroll
  for (int q = 0; q < AT_QPW; ++q) {
    int t = q0 + wid * AT_QPW + q;
    if (t >= p.N) continue;
    float inv = 1.f / lrow[q];
    long long row = (long long)s * p.N + t;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      int c = lane + 32 * i;
      if (c < d) st1(p.out, row * p.out_ld + head * d + c, p.oh, acc[q][i] * inv);
    }
    if (p.pesum) st1(p.pesum, (row * p.heads + head) * 32 + lane, p.ph_half, pacc[q] * inv);
  }
}

template <int MAXI>
static int launch_attn(const AttnP& p, cudaStream_t s) {
  size_t smem = sizeof(float) * ((AT_BQ + AT_BK) * (p.d + 4) + AT_BK * p.d);
  static bool attr_set[S2M2_MAX_DEVICES] = {false};
  const int di = current_device_index();
  if (!attr_set[di]) {
    cudaError_t e = set_func_attr(attention_simt_kernel<MAXI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    S2M2_CHECK(e == cudaSuccess, "attention_simt: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set[di] = true;
  }
  S2M2_CHECK(smem <= 200 * 1024, "attention_simt: head dim %d needs %zu bytes of shared memory (> 200 KB)", p.d, smem);
  S2M2_CHECK(p.S <= 65535 && p.heads <= 65535, "attention_simt: S = %d sequences exceed gridDim.z (65535); split the batch", p.S);
  dim3 grid(cdiv(p.N, AT_BQ), p.heads, p.S);
  launch_k(attention_simt_kernel<MAXI>, grid, AT_THREADS, smem, s, p);
  S2M2_LAUNCH_OK("attention_simt");
  return 0;
}

int attention_simt_launch(const AttnP& p, cudaStream_t s) {
  int maxi = (p.d + 31) / 32;
  if (maxi <= 1) return launch_attn<1>(p, s);
  if (maxi <= 2) return launch_attn<2>(p, s);
  if (maxi <= 4) return launch_attn<4>(p, s);
  if (maxi <= 8) return launch_attn<8>(p, s);
  if (maxi <= 12) return launch_attn<12>(p, s);
  set_error("attention: head dim %d > 384 not supported", p.d);
  return -1;
}

}  // namespace s2m2

using namespace s2m2;

extern "C" int64_t s2m2b200_attention_workspace_bytes(int32_t S, int32_t N, int32_t heads, int32_t d, int32_t dtype) {
  AttnP p{};
  p.S = S; p.N = N; p.heads = heads; p.d = d; p.C = heads * d; p.qh = p.oh = (dtype == S2M2_F16);
  alignas(16) static int dummy_aligned[4];
  p.qkv = dummy_aligned; p.out = dummy_aligned;   // alignment checks are on the real pointers at launch time
  p.qkv_ld = 3 * p.C; p.out_ld = p.C;
  if (!attention_tc_supported(p)) return 0;
  return (int64_t)attention_tc_workspace_bytes(p);
}

extern "C" int s2m2b200_attention(const S2m2Tensor* qkv, const S2m2Tensor* out, int32_t S, int32_t N, int32_t heads,
                                  int32_t kv_shift, const float* pe_x, const float* pe_y, int32_t ph, int32_t pw,
                                  const S2m2Tensor* pesum, void* work, int64_t work_bytes, void* stream) {
  S2M2_TRACE("attention", qkv, out, S, N, heads, kv_shift, pe_x, pe_y, ph, pw, pesum, work, work_bytes);
  const int C = out->C;
  S2M2_CHECK(qkv->C == 3 * C && C % heads == 0, "attention: qkv must have 3*C channels, C %% heads == 0");
  S2M2_CHECK(npix(*qkv) == (long long)S * N && npix(*out) == (long long)S * N, "attention: S*N != number of tokens");
  S2M2_CHECK(kv_shift >= 0 && kv_shift < S, "attention: bad kv_shift");
  AttnP p{};
  p.qkv = qkv->ptr; p.out = out->ptr; p.qkv_ld = qkv->ld; p.out_ld = out->ld;
  p.qh = qkv->dtype == S2M2_F16; p.oh = out->dtype == S2M2_F16;
  p.S = S; p.N = N; p.heads = heads; p.d = C / heads; p.C = C; p.kv_shift = kv_shift;
  p.scale = 1.0f / sqrtf((float)p.d);
  if (pesum && pesum->ptr) {
    S2M2_CHECK(pe_x && pe_y && ph * pw == N, "attention: PE needs tables and ph*pw == N");
    S2M2_CHECK(pesum->C == 32 && pesum->ld == 32 && npix(*pesum) == (long long)S * N * heads, "attention: pesum must be (S*N*heads, 32) dense");
    p.pesum = pesum->ptr; p.ph_half = pesum->dtype == S2M2_F16; p.pe_x = pe_x; p.pe_y = pe_y; p.pe_h = ph; p.pe_w = pw;
  }
  if (p.d % 4 == 0) {
    S2M2_CHECK(qkv->ld % 4 == 0 && aligned(qkv->ptr, 4 * esize(qkv->dtype)), "attention: qkv view must be 4-element aligned");
  }
  static const int impl = getenv("S2M2_ATTN_IMPL") ? atoi(getenv("S2M2_ATTN_IMPL")) : 0;   // 1 = CUDA-core, 2 = mma.sync, 0 = auto
  const bool force_simt = impl == 1;
  if (impl == 0 && work && attention_tc_supported(p) && work_bytes >= (int64_t)attention_tc_workspace_bytes(p))
    return attention_tc_launch(p, work, (cudaStream_t)stream);
  if (impl == 0 && attention2d_tc_supported(p, work, work_bytes)) return attention2d_tc_launch(p, work, (cudaStream_t)stream);
  if (!force_simt && attention_mma_supported(p)) return attention_mma_launch(p, (cudaStream_t)stream);
  return attention_simt_launch(p, (cudaStream_t)stream);
}
