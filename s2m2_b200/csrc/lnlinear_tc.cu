// Fused LayerNorm + Linear on tcgen05:   y = LN(x) W^T + b      (the pre-norm Q/K/V projection, attentions.py:119-130,150-161)
//
// A persistent CTA owns 128-token tiles; the x tile stays resident in shared memory (A operand) for all output-channel
// tiles, so LN(x) is never written to HBM and x is read once instead of once per 256 output channels:
//   TMA    x tile [128][C] fp16 -> shared memory (K-major, 128B swizzle)
//   LN     16 warps normalise the tile in place (fp32 statistics, no affine)
//   GEMM   for every 256-wide tile of output channels: tcgen05.mma into one of two TMEM accumulators, W streamed by TMA
//   EPI    TMEM -> +bias -> fp16 -> swizzled staging -> one cp.async.bulk.tensor store per warp; overlaps the next tile's MMAs
// warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..17 = LN / epilogue.
#include "tc_common.cuh"

namespace s2m2 {
using namespace tc;

constexpr int LNL_THREADS = 64 + 512;
constexpr int LNL_BN = 256;

struct LnlP {
  const float* bias;
  int M, N;
};
struct LnlMaps {
  CUtensorMap x;    // (tokens x C): box {64 ch, 128 tokens}
  CUtensorMap w;    // (N x C) K-major: box {64, 256}
  CUtensorMap out;  // (tokens x N): box {64 ch, 32 tokens}, 128B swizzle
};

__device__ __forceinline__ void lnl_tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <int C>
__global__ void __launch_bounds__(LNL_THREADS, 1) lnlinear_tc_kernel(const __grid_constant__ LnlMaps maps, const LnlP p) {
  constexpr int KC = C / 64;
  constexpr uint32_t X_BYTES = 128 * C * 2;
  constexpr uint32_t W_BYTES = LNL_BN * 128;
  constexpr uint32_t STG_BYTES = 16 * 4096;                     // 16 warps x [32 rows][64 ch] fp16
  constexpr int NST = (X_BYTES + STG_BYTES + 3 * W_BYTES <= 222 * 1024) ? 3 : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;
  uint8_t* sW = smem + X_BYTES;
  uint8_t* sStage = sW + NST * W_BYTES;
  uint64_t* bars = (uint64_t*)(sStage + STG_BYTES);            // xfull, xn_ready, a_free, tfull[2], tempty[2], wfull[NST], wempty[NST]
  uint32_t* tmem_slot = (uint32_t*)(bars + 7 + 2 * NST);
  const uint32_t xfull = smem_u32(bars), xn_ready = xfull + 8, a_free = xfull + 16, tfull0 = xfull + 24, tempty0 = xfull + 40;
  const uint32_t wfull0 = xfull + 56, wempty0 = wfull0 + 8 * NST;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + 127) / 128;
  const int n_tiles = (p.N + LNL_BN - 1) / LNL_BN;

  if (threadIdx.x == 0) {
    mbar_init(xfull, 1);
    mbar_init(xn_ready, 16);
    mbar_init(a_free, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 16); }
    for (int i = 0; i < NST; ++i) { mbar_init(wfull0 + 8 * i, 1); mbar_init(wempty0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * LNL_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t wit = 0, lt = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++lt) {
        if (lt > 0) mbar_wait(a_free, (lt - 1) & 1);            // every MMA of the previous token tile has read the A tile
        mbar_expect_tx(xfull, X_BYTES);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(&maps.x, smem_u32(sX + kc * 16384), xfull, kc * 64, tile * 128);
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int kb = 0; kb < KC; ++kb, ++wit) {
            const int st = wit % NST;
            mbar_wait(wempty0 + 8 * st, ((wit / NST) & 1) ^ 1);
            mbar_expect_tx(wfull0 + 8 * st, W_BYTES);
            tma_load_2d(&maps.w, smem_u32(sW + st * W_BYTES), wfull0 + 8 * st, kb * 64, nt * LNL_BN);
          }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(LNL_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t wit = 0, lt = 0, at = 0;   // W k-blocks, token tiles, accumulator uses
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++lt) {
        mbar_wait(xn_ready, lt & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int nt = 0; nt < n_tiles; ++nt, ++at) {
          const uint32_t acc = at & 1;
          mbar_wait(tempty0 + 8 * acc, ((at >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kb = 0; kb < KC; ++kb, ++wit) {
            const int st = wit % NST;
            mbar_wait(wfull0 + 8 * st, (wit / NST) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = make_sdesc(smem_u32(sX + kb * 16384)), bd = make_sdesc(smem_u32(sW + st * W_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + acc * LNL_BN, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
            umma_commit(wempty0 + 8 * st);
          }
          umma_commit(tfull0 + 8 * acc);
        }
        umma_commit(a_free);
      }
    }
  } else {
    // ================= LN + epilogue (16 warps) =================
    const int ew = warp - 2;
    const int q = warp & 3, cg = ew >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int lrow = ew * 8 + (lane >> 2), pt = lane & 3;
    uint8_t* stg = sStage + ew * 4096;
    const uint32_t stg_u32 = smem_u32(stg);
    uint32_t lt = 0, at = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++lt) {
      // ---- LayerNorm in place (two-pass fp32 statistics)
      mbar_wait(xfull, lt & 1);
      {
        uint8_t* rowp = sX + lrow * 128;
        float s = 0.f;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint4 v = *(const uint4*)(rowp + kc * 16384 + (((2 * pt + u) ^ (lrow & 7)) << 4));
            const __half2* h = (const __half2*)&v;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); s += f.x + f.y; }
          }
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float mean = s * (1.0f / C);
        float qv = 0.f;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint4 v = *(const uint4*)(rowp + kc * 16384 + (((2 * pt + u) ^ (lrow & 7)) << 4));
            const __half2* h = (const __half2*)&v;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); qv += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean); }
          }
        qv += __shfl_xor_sync(0xffffffffu, qv, 1); qv += __shfl_xor_sync(0xffffffffu, qv, 2);
        const float rstd = rsqrtf(qv * (1.0f / C) + 1e-5f);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint4* pv4 = (uint4*)(rowp + kc * 16384 + (((2 * pt + u) ^ (lrow & 7)) << 4));
            uint4 v = *pv4;
            __half2* h = (__half2*)&v;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(h[i]);
              h[i] = __floats2half2_rn((f.x - mean) * rstd, (f.y - mean) * rstd);
            }
            *pv4 = v;
          }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(xn_ready);

      // ---- epilogues: one per 256-wide output-channel tile
      for (int nt = 0; nt < n_tiles; ++nt, ++at) {
        const uint32_t acc = at & 1;
        const int n0 = nt * LNL_BN + cg * 64;                     // this warp's 64 output channels
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
        __syncwarp();
        mbar_wait(tfull0 + 8 * acc, (at >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int n = n0 + c * 16;
          if (n >= p.N) break;
          uint32_t rr[16];
          tmem_ld16(trow + acc * LNL_BN + cg * 64 + c * 16, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float v[16];
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bb = __ldg((const float4*)(p.bias + n + i));
              v[i] = __uint_as_float(rr[i]) + bb.x; v[i + 1] = __uint_as_float(rr[i + 1]) + bb.y;
              v[i + 2] = __uint_as_float(rr[i + 2]) + bb.z; v[i + 3] = __uint_as_float(rr[i + 3]) + bb.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
          }
          __half2 h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t lin = (uint32_t)lane * 128 + (uint32_t)(c * 2 + u) * 16;
            *(uint4*)(stg + (lin ^ (((lin >> 7) & 7u) << 4))) = *(uint4*)&h[4 * u];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tempty0 + 8 * acc);
          if (n0 < p.N) {
            lnl_tma_store_2d(&maps.out, stg_u32, n0, tile * 128 + q * 32);   // clipped at N and M by the tensor map
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * LNL_BN));
  }
}

template <int C>
static int launch_lnl(const LnlMaps& maps, const LnlP& p, cudaStream_t s) {
  constexpr size_t XB = 128 * C * 2, WB = LNL_BN * 128, SB = 16 * 4096;
  constexpr int NST = (XB + SB + 3 * WB <= 222 * 1024) ? 3 : 2;
  constexpr size_t smem = XB + NST * WB + SB + (7 + 2 * NST) * 8 + 16 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr[S2M2_MAX_DEVICES] = {false};
  static int sms_of[S2M2_MAX_DEVICES] = {0};
  const int di = current_device_index();
  int& sms = sms_of[di];
  if (!attr[di]) {
    cudaError_t e = set_func_attr(lnlinear_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    S2M2_CHECK(e == cudaSuccess, "lnlinear_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    attr[di] = true;
  }
  const int tiles = (p.M + 127) / 128;
  launch_k(lnlinear_tc_kernel<C>, tiles < sms ? tiles : sms, LNL_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("lnlinear_tc");
  return 0;
}

}  // namespace s2m2

using namespace s2m2;

// returns 0 = launched, 1 = not applicable (caller runs layernorm + conv2d), < 0 = error
extern "C" int s2m2b200_ln_linear(const S2m2Tensor* x, const S2m2Tensor* y, const void* w, const float* bias, int32_t w_dtype,
                                  void* stream) {
  S2M2_TRACE("ln_linear", x, y, w, bias, w_dtype);
  const int C = x->C, N = y->C;
  if (x->dtype != S2M2_F16 || y->dtype != S2M2_F16 || w_dtype != S2M2_F16) return 1;
  if (!(C == 128 || C == 192 || C == 256)) return 1;
  if (N % 64 || x->ld % 8 || y->ld % 8 || !aligned(x->ptr, 16) || !aligned(y->ptr, 16) || !aligned(w, 16) || (bias && !aligned(bias, 16)))
    return 1;
  S2M2_CHECK(npix(*x) == npix(*y), "ln_linear: x / y token count mismatch");
  if (!tc::device_ok()) return 1;
  LnlP p{};
  p.bias = bias; p.M = (int)npix(*x); p.N = N;
  LnlMaps maps;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)p.M}, strides[1] = {(cuuint64_t)x->ld * 2};
    cuuint32_t box[2] = {64, 128};
    CUresult r = tc::encode_fn()(&maps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "ln_linear: cuTensorMapEncodeTiled(x) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, LNL_BN};
    CUresult r = tc::encode_fn()(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "ln_linear: cuTensorMapEncodeTiled(w) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)p.M}, strides[1] = {(cuuint64_t)y->ld * 2};
    cuuint32_t box[2] = {64, 32};
    CUresult r = tc::encode_fn()(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "ln_linear: cuTensorMapEncodeTiled(out) failed: %d", (int)r);
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (C) {
    case 128: return launch_lnl<128>(maps, p, s);
    case 192: return launch_lnl<192>(maps, p, s);
    default: return launch_lnl<256>(maps, p, s);
  }
}
