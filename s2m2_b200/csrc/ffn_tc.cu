// Fused pre-norm FFN on tcgen05:   y = x + W2 . GELU(W0 . LN(x) + b0) + b2      (attentions.py:229-250, dim_expansion = 1)
//
// A persistent CTA owns 128-token tiles.  Per tile the activations never leave the SM:
//   TMA    x tile [128][C] fp16 -> shared memory in the K-major 128B-swizzled A-operand layout
//   LN     16 warps normalise the tile in place (fp32 statistics, no affine)            -> A operand of GEMM 1
//   GEMM1  H = LN(x) W0^T   tcgen05.mma, W0 streamed through a TMA ring, accumulators in TMEM (C columns)
//   EPI1   TMEM -> +b0 -> GELU -> fp16, written back over the tile in the same A-operand layout   -> A operand of GEMM 2
//   GEMM2  Y = H W2^T       into the same TMEM columns
//   EPI2   TMEM -> +b2 + x (residual re-read from global, prefetched) -> fp16 -> global (full-sector stores)
// This replaces three launches (LayerNorm, two linears) and the two HBM round trips of LN(x) and H.
// warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..17 = LN / epilogues.
#include <stdlib.h>

#include "tc_common.cuh"

namespace s2m2 {
using namespace tc;

constexpr int FFN_THREADS = 64 + 512;

struct FfnP {
  const float *b0, *b2;
  const void* x;
  void* y;
  long long x_ld, y_ld;
  int M, tmem_cols;
};
struct FfnMaps {
  CUtensorMap x;    // (tokens x C): box {64 ch, 128 tokens}
  CUtensorMap w0;   // (C x C) K-major: box {64, BNW}
  CUtensorMap w2;
};

template <int C>
__global__ void __launch_bounds__(FFN_THREADS, 1) ffn_tc_kernel(const __grid_constant__ FfnMaps maps, const FfnP p) {
  constexpr int KC = C / 64;                                   // 64-channel chunks = k-blocks
  constexpr int BNW = (C <= 256) ? C : ((C % 256 == 0) ? 256 : 128);   // MMA N per instruction
  constexpr int NT = C / BNW;
  constexpr uint32_t X_BYTES = 128 * C * 2;
  constexpr uint32_t W_BYTES = BNW * 128;
  // the residual x is kept in a second shared-memory copy of the tile when it fits (C <= 384); otherwise it is re-read from global
  constexpr bool RES_SMEM = (2 * X_BYTES + 2 * W_BYTES <= 222 * 1024);
  constexpr uint32_t XR_BYTES = RES_SMEM ? 2 * X_BYTES : X_BYTES;
  constexpr int NST = (XR_BYTES + 4 * W_BYTES <= 222 * 1024) ? 4 : ((XR_BYTES + 3 * W_BYTES <= 222 * 1024) ? 3 : 2);
  constexpr int CG = C / 4;                                     // columns per epilogue warp group
  static_assert(CG % 16 == 0 && C % 64 == 0, "channel count");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                                           // x -> LN(x) -> H, KC chunks of [128 rows][128 B]
  uint8_t* sR = smem + X_BYTES;                                 // RES_SMEM: untouched copy of x for the residual
  uint8_t* sW = smem + XR_BYTES;                                // NST x [BNW rows][128 B]
  uint64_t* bars = (uint64_t*)(sW + NST * W_BYTES);             // xfull, xn_ready, h_full, h_ready, y_full, r_free, wfull[NST], wempty[NST]
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 + 2 * NST);
  const uint32_t xfull = smem_u32(bars), xn_ready = xfull + 8, h_full = xfull + 16, h_ready = xfull + 24, y_full = xfull + 32;
  const uint32_t r_free = xfull + 40;                           // epilogue 2 has finished reading the residual copy
  const uint32_t wfull0 = xfull + 48, wempty0 = wfull0 + 8 * NST;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(xfull, 1);
    mbar_init(xn_ready, 16);
    mbar_init(h_full, 1);
    mbar_init(h_ready, 16);
    mbar_init(y_full, 1);
    mbar_init(r_free, 16);
    for (int i = 0; i < NST; ++i) { mbar_init(wfull0 + 8 * i, 1); mbar_init(wempty0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
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
        if (lt > 0) {
          mbar_wait(y_full, (lt - 1) & 1);                       // GEMM 2 of the previous tile no longer reads the tile buffer
          if (RES_SMEM) mbar_wait(r_free, (lt - 1) & 1);         // ... and its epilogue no longer reads the residual copy
        }
        mbar_expect_tx(xfull, XR_BYTES);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(&maps.x, smem_u32(sX + kc * 16384), xfull, kc * 64, tile * 128);
        if (RES_SMEM) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) tma_load_2d(&maps.x, smem_u32(sR + kc * 16384), xfull, kc * 64, tile * 128);
        }
        for (int g = 0; g < 2; ++g)
          for (int nt = 0; nt < NT; ++nt)
            for (int kb = 0; kb < KC; ++kb, ++wit) {
              const int st = wit % NST;
              mbar_wait(wempty0 + 8 * st, ((wit / NST) & 1) ^ 1);
              mbar_expect_tx(wfull0 + 8 * st, W_BYTES);
              tma_load_2d(g == 0 ? &maps.w0 : &maps.w2, smem_u32(sW + st * W_BYTES), wfull0 + 8 * st, kb * 64, nt * BNW);
            }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BNW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t wit = 0, lt = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++lt) {
        for (int g = 0; g < 2; ++g) {
          mbar_wait(g == 0 ? xn_ready : h_ready, lt & 1);        // A operand is in shared memory, TMEM columns are free
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int nt = 0; nt < NT; ++nt)
            for (int kb = 0; kb < KC; ++kb, ++wit) {
              const int st = wit % NST;
              mbar_wait(wfull0 + 8 * st, (wit / NST) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint64_t ad = make_sdesc(smem_u32(sX + kb * 16384)), bd = make_sdesc(smem_u32(sW + st * W_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tmem_base + nt * BNW, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
              umma_commit(wempty0 + 8 * st);
            }
          umma_commit(g == 0 ? h_full : y_full);
        }
      }
    }
  } else {
    // ================= LN + epilogues (16 warps) =================
    const int ew = warp - 2;
    const int q = warp & 3, cg = ew >> 2;                        // TMEM lane quarter, column group
    const int r = q * 32 + lane;                                  // epilogue row
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // LayerNorm mapping: warp ew owns rows 8*ew .. 8*ew+7; 4 lanes per row, lane part pt covers 16-byte units {2pt, 2pt+1} of each chunk
    const int lrow = ew * 8 + (lane >> 2), pt = lane & 3;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++lt) {
      const uint32_t ph = lt & 1;
      const long long tok = (long long)tile * 128 + r;
      const bool valid = tok < p.M;
      constexpr int NCH = CG / 16;
      // ---- LayerNorm in place (two-pass statistics in fp32, like elem.cu:layernorm_kernel)
      mbar_wait(xfull, ph);
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
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // (epilogue 2 of the previous tile read TMEM before this point)
      __syncwarp();
      if (lane == 0) mbar_arrive(xn_ready);

      // ---- epilogue 1: H = GELU(acc + b0) -> fp16 -> A-operand layout, over the tile buffer
      mbar_wait(h_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = cg * CG + c * 16;
        uint32_t rr[16];
        tmem_ld16(trow + col, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 bb = __ldg((const float4*)(p.b0 + col + i));
          v[i] = __uint_as_float(rr[i]) + bb.x; v[i + 1] = __uint_as_float(rr[i + 1]) + bb.y;
          v[i + 2] = __uint_as_float(rr[i + 2]) + bb.z; v[i + 3] = __uint_as_float(rr[i + 3]) + bb.w;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);
        __half2 h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        uint8_t* blk = sX + (col >> 6) * 16384 + r * 128;
        const int u0 = (col & 63) >> 3;
        *(uint4*)(blk + (((u0) ^ (r & 7)) << 4)) = *(uint4*)&h[0];
        *(uint4*)(blk + (((u0 + 1) ^ (r & 7)) << 4)) = *(uint4*)&h[4];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);

      // ---- epilogue 2: y = acc + b2 + x
      // residual: from the shared-memory copy when it fits, else re-read from global per chunk (full-sector lane-pair reads)
      const bool odd = lane & 1;
      const bool pvalid = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
      const long long ptok = __shfl_xor_sync(0xffffffffu, tok, 1);
      const long long row_e = odd ? ptok : tok, row_o = odd ? tok : ptok;
      const bool ve = odd ? pvalid : valid, vo = odd ? valid : pvalid;
      mbar_wait(y_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long off0 = tok * p.y_ld + cg * CG;
      const long long poff0 = __shfl_xor_sync(0xffffffffu, off0, 1);
      __half* yb = (__half*)p.y;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = cg * CG + c * 16;
        uint32_t rr[16];
        tmem_ld16(trow + col, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 bb = __ldg((const float4*)(p.b2 + col + i));
          v[i] = __uint_as_float(rr[i]) + bb.x; v[i + 1] = __uint_as_float(rr[i + 1]) + bb.y;
          v[i + 2] = __uint_as_float(rr[i + 2]) + bb.z; v[i + 3] = __uint_as_float(rr[i + 3]) + bb.w;
        }
        uint4 r2[2];
        if (RES_SMEM) {   // this thread's 16 columns of x from the untouched shared-memory copy (same swizzled layout as the tile)
          const uint8_t* rblk = sR + (col >> 6) * 16384 + r * 128;
          const int u0 = (col & 63) >> 3;
          r2[0] = *(const uint4*)(rblk + (((u0) ^ (r & 7)) << 4));
          r2[1] = *(const uint4*)(rblk + (((u0 + 1) ^ (r & 7)) << 4));
        } else {
          const __half* rb = (const __half*)p.x + col + (odd ? 8 : 0);
          uint4 a = make_uint4(0, 0, 0, 0), b = a;
          if (ve) a = *(const uint4*)(rb + row_e * p.x_ld);   // plain loads: y may alias x (in-place residual stream)
          if (vo) b = *(const uint4*)(rb + row_o * p.x_ld);
          const uint4 sendr = odd ? a : b;
          uint4 recvr;
          recvr.x = __shfl_xor_sync(0xffffffffu, sendr.x, 1); recvr.y = __shfl_xor_sync(0xffffffffu, sendr.y, 1);
          recvr.z = __shfl_xor_sync(0xffffffffu, sendr.z, 1); recvr.w = __shfl_xor_sync(0xffffffffu, sendr.w, 1);
          r2[0] = odd ? recvr : a;
          r2[1] = odd ? b : recvr;
        }
        const __half2* hr = (const __half2*)&r2[0];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 f = __half22float2(hr[i]);
          v[2 * i] += f.x; v[2 * i + 1] += f.y;
        }
        __half2 h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const uint4 lo = *(uint4*)&h[0], hi = *(uint4*)&h[4];
        const uint4 send = odd ? lo : hi;
        uint4 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
        if (odd ? pvalid : valid) *(uint4*)(yb + (odd ? poff0 + c * 16 + 8 : off0 + c * 16)) = odd ? recv : lo;
        if (odd ? valid : pvalid) *(uint4*)(yb + (odd ? off0 + c * 16 + 8 : poff0 + c * 16)) = odd ? hi : recv;
      }
      if (RES_SMEM) {
        __syncwarp();
        if (lane == 0) mbar_arrive(r_free);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}


// =====================================================================================================
// v2 (C <= 256): two token tiles in flight per CTA.  The v1 kernel above runs its per-tile chain
// (TMA -> LN -> GEMM1 -> GELU -> GEMM2 -> residual) strictly in sequence, so the tensor pipe idles during the epilogues and the
// 16 epilogue warps idle during the GEMMs (16 us per tile at C = 256, of which 3.4 us are MMAs).  Here tiles alternate between
// two shared-memory tile buffers and two TMEM accumulators: the epilogue warps work on tile A while the MMAs of tile B run,
//   epilogue warps:  LN(A) LN(B) | EPI1(A) EPI1(B) | EPI2(A) EPI2(B) | LN(A') ...
//   MMA thread:            G1(A) G1(B)       G2(A)   G2(B)                  G1(A') ...
// and the next pair's x tiles are fetched as soon as GEMM 2 has released a buffer.  The same kernel serves the 1x1 branch of
// ConvBlock2D (attentions.py:273-275: Conv1x1 -> ReLU -> Conv1x1, summed with the 3x3 branch): LN = false, ACT = ReLU, the
// residual read from a third tensor.
//   y = [res +] W2 . ACT(W0 . (LN ? LayerNorm(x) : x) + b0) + b2
// =====================================================================================================
struct Mlp2P {
  const float *b0, *b2;
  const void* res;     // residual rows (fp16), may alias y (and x); nullptr = none
  void* y;
  long long res_ld, y_ld;
  int M, tmem_cols;
};

template <int C, bool LN, int ACT>
__global__ void __launch_bounds__(FFN_THREADS, 1) mlp2_tc_kernel(const __grid_constant__ FfnMaps maps, const Mlp2P p) {
  constexpr int KC = C / 64;
  constexpr uint32_t X_BYTES = 128 * C * 2;
  constexpr uint32_t W_BYTES = C * 128;                          // one k-block of the (C x C) weights: C rows x 128 B
  constexpr int NST_FIT = (224 * 1024 - 2 * (int)X_BYTES) / (int)W_BYTES;
  constexpr int NST = NST_FIT > 6 ? 6 : NST_FIT;
  static_assert(NST >= 3, "weight ring too shallow");
  constexpr int CG = C / 4;
  constexpr int NCH = CG / 16;
  static_assert(CG % 16 == 0 && C % 64 == 0 && C <= 256, "channel count");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX0 = smem;                                           // tile buffer of slot 0 / 1: x -> LN(x) -> H
  uint8_t* sW = smem + 2 * X_BYTES;
  uint64_t* bars = (uint64_t*)(sW + NST * W_BYTES);              // per slot: xfull, xn_ready, h_full, h_ready, y_full; wfull[NST], wempty[NST]
  uint32_t* tmem_slot = (uint32_t*)(bars + 10 + 2 * NST);
  const uint32_t bar0 = smem_u32(bars);
  auto xfull = [&](int s) { return bar0 + 8 * (5 * s); };
  auto xn_ready = [&](int s) { return bar0 + 8 * (5 * s + 1); };
  auto h_full = [&](int s) { return bar0 + 8 * (5 * s + 2); };
  auto h_ready = [&](int s) { return bar0 + 8 * (5 * s + 3); };
  auto y_full = [&](int s) { return bar0 + 8 * (5 * s + 4); };
  const uint32_t wfull0 = bar0 + 80, wempty0 = wfull0 + 8 * NST;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + 127) / 128;
  // this CTA's tiles: blockIdx.x + i * gridDim.x, i = 0 .. n_local-1; tile i lives in slot i & 1, pairs (2j, 2j+1) travel together
  const int n_local = (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xn_ready(s), 16);
      mbar_init(h_full(s), 1);
      mbar_init(h_ready(s), 16);
      mbar_init(y_full(s), 1);
    }
    for (int i = 0; i < NST; ++i) { mbar_init(wfull0 + 8 * i, 1); mbar_init(wempty0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // ================= TMA producer: x tiles of a pair, then the pair's weights in the order the MMA thread consumes them =================
    if (lane == 0) {
      uint32_t wit = 0;
      for (int i0 = 0; i0 < n_local; i0 += 2) {
        const int np = (i0 + 1 < n_local) ? 2 : 1;
        for (int s = 0; s < np; ++s) {
          const int i = i0 + s;
          if (i >= 2) mbar_wait(y_full(s), ((i - 2) >> 1) & 1);   // GEMM 2 of the slot's previous tile has finished reading the buffer
          mbar_expect_tx(xfull(s), X_BYTES);
          const int tile = (int)blockIdx.x + i * (int)gridDim.x;
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) tma_load_2d(&maps.x, smem_u32(sX0 + s * X_BYTES + kc * 16384), xfull(s), kc * 64, tile * 128);
        }
        for (int g = 0; g < 2; ++g)
          for (int s = 0; s < np; ++s)
            for (int kb = 0; kb < KC; ++kb, ++wit) {
              const int st = wit % NST;
              mbar_wait(wempty0 + 8 * st, ((wit / NST) & 1) ^ 1);
              mbar_expect_tx(wfull0 + 8 * st, W_BYTES);
              tma_load_2d(g == 0 ? &maps.w0 : &maps.w2, smem_u32(sW + st * W_BYTES), wfull0 + 8 * st, kb * 64, 0);
            }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t wit = 0;
      for (int i0 = 0; i0 < n_local; i0 += 2) {
        const int np = (i0 + 1 < n_local) ? 2 : 1;
        const uint32_t ph = (i0 >> 1) & 1;
        for (int g = 0; g < 2; ++g)
          for (int s = 0; s < np; ++s) {
            mbar_wait(g == 0 ? xn_ready(s) : h_ready(s), ph);      // A operand in shared memory, accumulator columns free
            if (g == 0) mbar_wait(xfull(s), ph);                   // (the TMA's bytes, observed by this thread as well)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < KC; ++kb, ++wit) {
              const int st = wit % NST;
              mbar_wait(wfull0 + 8 * st, (wit / NST) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint64_t ad = make_sdesc(smem_u32(sX0 + s * X_BYTES + kb * 16384)), bd = make_sdesc(smem_u32(sW + st * W_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tmem_base + s * C, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
              umma_commit(wempty0 + 8 * st);
            }
            umma_commit(g == 0 ? h_full(s) : y_full(s));
          }
      }
    }
  } else {
    // ================= LN + epilogues (16 warps) =================
    const int ew = warp - 2;
    const int q = warp & 3, cg = ew >> 2;                        // TMEM lane quarter, column group
    const int r = q * 32 + lane;                                  // epilogue row
    const int lrow = ew * 8 + (lane >> 2), pt = lane & 3;        // LayerNorm mapping: 4 lanes per row
    const bool odd = lane & 1;
    for (int i0 = 0; i0 < n_local; i0 += 2) {
      const int np = (i0 + 1 < n_local) ? 2 : 1;
      const uint32_t ph = (i0 >> 1) & 1;
      // ---- LayerNorm in place (or nothing): the tile becomes the A operand of GEMM 1
      for (int s = 0; s < np; ++s) {
        mbar_wait(xfull(s), ph);
        if constexpr (LN) {
          uint8_t* rowp = sX0 + s * X_BYTES + lrow * 128;
          float sum = 0.f;
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const uint4 v = *(const uint4*)(rowp + kc * 16384 + (((2 * pt + u) ^ (lrow & 7)) << 4));
              const __half2* h = (const __half2*)&v;
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); sum += f.x + f.y; }
            }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1); sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          const float mean = sum * (1.0f / C);
          float qv = 0.f;
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const uint4 v = *(const uint4*)(rowp + kc * 16384 + (((2 * pt + u) ^ (lrow & 7)) << 4));
              const __half2* h = (const __half2*)&v;
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); qv += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean); }
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
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                h[j] = __floats2half2_rn((f.x - mean) * rstd, (f.y - mean) * rstd);
              }
              *pv4 = v;
            }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // epilogue 2 of the slot's previous tile read TMEM before this point
        __syncwarp();
        if (lane == 0) mbar_arrive(xn_ready(s));
      }
      // ---- epilogue 1: H = ACT(acc + b0) -> fp16 -> A-operand layout, over the tile buffer
      for (int s = 0; s < np; ++s) {
        const uint32_t trow = tmem_base + s * C + ((uint32_t)(q * 32) << 16);
        uint8_t* sX = sX0 + s * X_BYTES;
        mbar_wait(h_full(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t rbuf[2][16];
        tmem_ld16(trow + cg * CG, rbuf[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col = cg * CG + c * 16;
          uint32_t* rr = rbuf[c & 1];
          tmem_ld_wait16(rr);
          if (c + 1 < NCH) tmem_ld16(trow + col + 16, rbuf[(c + 1) & 1]);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = __ldg((const float4*)(p.b0 + col + j));
            v[j] = __uint_as_float(rr[j]) + bb.x; v[j + 1] = __uint_as_float(rr[j + 1]) + bb.y;
            v[j + 2] = __uint_as_float(rr[j + 2]) + bb.z; v[j + 3] = __uint_as_float(rr[j + 3]) + bb.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (ACT == S2M2_ACT_GELU) ? gelu_fast(v[j]) : fmaxf(v[j], 0.0f);
          __half2 h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          uint8_t* blk = sX + (col >> 6) * 16384 + r * 128;
          const int u0 = (col & 63) >> 3;
          *(uint4*)(blk + (((u0) ^ (r & 7)) << 4)) = *(uint4*)&h[0];
          *(uint4*)(blk + (((u0 + 1) ^ (r & 7)) << 4)) = *(uint4*)&h[4];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(h_ready(s));
      }
      // ---- epilogue 2: y = acc + b2 [+ res]   (residual rows re-read from global / L2 with full-sector lane-pair loads)
      for (int s = 0; s < np; ++s) {
        const uint32_t trow = tmem_base + s * C + ((uint32_t)(q * 32) << 16);
        const int tile = (int)blockIdx.x + (i0 + s) * (int)gridDim.x;
        const long long tok = (long long)tile * 128 + r;
        const bool valid = tok < p.M;
        const bool pvalid = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
        const long long ptok = __shfl_xor_sync(0xffffffffu, tok, 1);
        const long long row_e = odd ? ptok : tok, row_o = odd ? tok : ptok;
        const bool ve = odd ? pvalid : valid, vo = odd ? valid : pvalid;
        // the residual does not depend on the accumulator: fetch it before waiting for GEMM 2
        uint4 ra[NCH], rb[NCH];
        if (p.res) {
          const __half* rbase = (const __half*)p.res + cg * CG + (odd ? 8 : 0);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            ra[c] = make_uint4(0, 0, 0, 0); rb[c] = ra[c];
            if (ve) ra[c] = *(const uint4*)(rbase + c * 16 + row_e * p.res_ld);   // plain loads: y may alias the residual
            if (vo) rb[c] = *(const uint4*)(rbase + c * 16 + row_o * p.res_ld);
          }
        }
        mbar_wait(y_full(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long off0 = tok * p.y_ld + cg * CG;
        const long long poff0 = __shfl_xor_sync(0xffffffffu, off0, 1);
        __half* yb = (__half*)p.y;
        uint32_t rbuf[2][16];
        tmem_ld16(trow + cg * CG, rbuf[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col = cg * CG + c * 16;
          uint32_t* rr = rbuf[c & 1];
          tmem_ld_wait16(rr);
          if (c + 1 < NCH) tmem_ld16(trow + col + 16, rbuf[(c + 1) & 1]);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = __ldg((const float4*)(p.b2 + col + j));
            v[j] = __uint_as_float(rr[j]) + bb.x; v[j + 1] = __uint_as_float(rr[j + 1]) + bb.y;
            v[j + 2] = __uint_as_float(rr[j + 2]) + bb.z; v[j + 3] = __uint_as_float(rr[j + 3]) + bb.w;
          }
          if (p.res) {
            const uint4 sendr = odd ? ra[c] : rb[c];
            uint4 recvr;
            recvr.x = __shfl_xor_sync(0xffffffffu, sendr.x, 1); recvr.y = __shfl_xor_sync(0xffffffffu, sendr.y, 1);
            recvr.z = __shfl_xor_sync(0xffffffffu, sendr.z, 1); recvr.w = __shfl_xor_sync(0xffffffffu, sendr.w, 1);
            uint4 r2[2];
            r2[0] = odd ? recvr : ra[c];
            r2[1] = odd ? rb[c] : recvr;
            const __half2* hr = (const __half2*)&r2[0];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 f = __half22float2(hr[j]);
              v[2 * j] += f.x; v[2 * j + 1] += f.y;
            }
          }
          __half2 h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          const uint4 lo = *(uint4*)&h[0], hi = *(uint4*)&h[4];
          const uint4 send = odd ? lo : hi;
          uint4 recv;
          recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
          recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
          if (odd ? pvalid : valid) *(uint4*)(yb + (odd ? poff0 + c * 16 + 8 : off0 + c * 16)) = odd ? recv : lo;
          if (odd ? valid : pvalid) *(uint4*)(yb + (odd ? off0 + c * 16 + 8 : poff0 + c * 16)) = odd ? hi : recv;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

template <int C, bool LN, int ACT>
static int launch_mlp2(const FfnMaps& maps, const Mlp2P& p, cudaStream_t s) {
  constexpr size_t XB = 128 * C * 2, WB = C * 128;
  constexpr int NST_FIT = (224 * 1024 - 2 * (int)XB) / (int)WB;
  constexpr int NST = NST_FIT > 6 ? 6 : NST_FIT;
  constexpr size_t smem = 2 * XB + NST * WB + (10 + 2 * NST) * 8 + 16 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {0};
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms_of[dev]) {   // per device: the attribute belongs to the (function, device) pair
    cudaError_t e = set_func_attr(mlp2_tc_kernel<C, LN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    S2M2_CHECK(e == cudaSuccess, "mlp2_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sms_of[dev] = sms > 0 ? sms : 1;
  }
  const int tiles = (p.M + 127) / 128;
  const int sms = sms_of[dev];
  // an even number of tiles per CTA keeps both slots busy: with fewer than two tiles per SM use half the CTAs
  int grid = tiles < sms ? tiles : sms;
  if (tiles < 2 * sms && tiles > 1) grid = (tiles + 1) / 2;
  launch_k(mlp2_tc_kernel<C, LN, ACT>, grid, FFN_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("mlp2_tc");
  return 0;
}

template <int C>
static int launch_ffn(const FfnMaps& maps, const FfnP& p, cudaStream_t s) {
  constexpr int BNW = (C <= 256) ? C : ((C % 256 == 0) ? 256 : 128);
  constexpr size_t XB = 128 * C * 2, WB = BNW * 128;
  constexpr bool RES_SMEM = (2 * XB + 2 * WB <= 222 * 1024);
  constexpr size_t XR = RES_SMEM ? 2 * XB : XB;
  constexpr int NST = (XR + 4 * WB <= 222 * 1024) ? 4 : ((XR + 3 * WB <= 222 * 1024) ? 3 : 2);
  constexpr size_t smem = XR + NST * WB + (6 + 2 * NST) * 8 + 16 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr[S2M2_MAX_DEVICES] = {false};
  static int sms_of[S2M2_MAX_DEVICES] = {0};
  const int di = current_device_index();
  int& sms = sms_of[di];
  if (!attr[di]) {
    cudaError_t e = set_func_attr(ffn_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    S2M2_CHECK(e == cudaSuccess, "ffn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    attr[di] = true;
  }
  const int tiles = (p.M + 127) / 128;
  launch_k(ffn_tc_kernel<C>, tiles < sms ? tiles : sms, FFN_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("ffn_tc");
  return 0;
}

}  // namespace s2m2

using namespace s2m2;

static int encode_ffn_maps(FfnMaps& maps, const S2m2Tensor* x, int M, int C, int BNW, const void* w0, const void* w2, const char* who) {
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M}, strides[1] = {(cuuint64_t)x->ld * 2};
    cuuint32_t box[2] = {64, 128};
    CUresult r = tc::encode_fn()(&maps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(x) failed: %d", who, (int)r);
  }
  const void* ws[2] = {w0, w2};
  CUtensorMap* wm[2] = {&maps.w0, &maps.w2};
  for (int g = 0; g < 2; ++g) {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)C}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BNW};
    CUresult r = tc::encode_fn()(wm[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ws[g]), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(w) failed: %d", who, (int)r);
  }
  return 0;
}

// y = [res +] W2 . act(W0 . (ln ? LayerNorm(x) : x) + b0) + b2 with the two-tile kernel; 0 = launched, 1 = not applicable, < 0 = error
static int mlp2_dispatch(const S2m2Tensor* x, const S2m2Tensor* res, const S2m2Tensor* y, const void* w0, const float* b0, const void* w2,
                         const float* b2, int32_t w_dtype, int ln, int act, cudaStream_t s) {
  const int C = x->C;
  if (x->dtype != S2M2_F16 || y->dtype != S2M2_F16 || w_dtype != S2M2_F16 || (res && res->ptr && res->dtype != S2M2_F16)) return 1;
  if (!(C == 128 || C == 192 || C == 256)) return 1;
  if (!((ln && act == S2M2_ACT_GELU) || (!ln && act == S2M2_ACT_RELU))) return 1;
  if (x->ld % 8 || y->ld % 8 || !aligned(x->ptr, 16) || !aligned(y->ptr, 16) || !aligned(w0, 16) || !aligned(w2, 16) ||
      !aligned(b0, 16) || !aligned(b2, 16) || !b0 || !b2)
    return 1;
  if (res && res->ptr && (res->ld % 8 || !aligned(res->ptr, 16))) return 1;
  S2M2_CHECK(y->C == C && npix(*x) == npix(*y), "mlp2: x / y shape mismatch");
  S2M2_CHECK(!(res && res->ptr) || (res->C == C && npix(*res) == npix(*x)), "mlp2: residual shape mismatch");
  if (!tc::device_ok()) return 1;
  Mlp2P p{};
  p.b0 = b0; p.b2 = b2; p.y = y->ptr; p.y_ld = y->ld;
  p.res = (res && res->ptr) ? res->ptr : nullptr; p.res_ld = p.res ? res->ld : 0;
  p.M = (int)npix(*x);
  p.tmem_cols = C <= 128 ? 256 : 512;
  FfnMaps maps;
  if (int r = encode_ffn_maps(maps, x, p.M, C, C, w0, w2, "mlp2")) return r;
  if (ln) {
    switch (C) {
      case 128: return launch_mlp2<128, true, S2M2_ACT_GELU>(maps, p, s);
      case 192: return launch_mlp2<192, true, S2M2_ACT_GELU>(maps, p, s);
      default: return launch_mlp2<256, true, S2M2_ACT_GELU>(maps, p, s);
    }
  }
  switch (C) {
    case 128: return launch_mlp2<128, false, S2M2_ACT_RELU>(maps, p, s);
    case 192: return launch_mlp2<192, false, S2M2_ACT_RELU>(maps, p, s);
    default: return launch_mlp2<256, false, S2M2_ACT_RELU>(maps, p, s);
  }
}

// returns 0 = launched, 1 = shape / dtype not supported by the fused kernel (caller runs LayerNorm + two conv2d), < 0 = error
extern "C" int s2m2b200_ffn(const S2m2Tensor* x, const S2m2Tensor* y, const void* w0, const float* b0, const void* w2, const float* b2,
                            int32_t w_dtype, void* stream) {
  S2M2_TRACE("ffn", x, y, w0, b0, w2, b2, w_dtype);
  const int C = x->C;
  static const int force_v1 = getenv("S2M2_FFN_V1") ? atoi(getenv("S2M2_FFN_V1")) : 0;
  if (!force_v1 && C <= 256) {   // two tiles in flight; the residual is x itself
    const int r = mlp2_dispatch(x, x, y, w0, b0, w2, b2, w_dtype, 1, S2M2_ACT_GELU, (cudaStream_t)stream);
    if (r != 1) return r;
  }
  if (x->dtype != S2M2_F16 || y->dtype != S2M2_F16 || w_dtype != S2M2_F16) return 1;
  if (!(C == 128 || C == 192 || C == 256 || C == 384 || C == 512)) return 1;
  if (x->ld % 8 || y->ld % 8 || !aligned(x->ptr, 16) || !aligned(y->ptr, 16) || !aligned(w0, 16) || !aligned(w2, 16) ||
      !aligned(b0, 16) || !aligned(b2, 16) || !b0 || !b2)
    return 1;
  S2M2_CHECK(y->C == C && npix(*x) == npix(*y), "ffn: x / y shape mismatch");
  if (!tc::device_ok()) return 1;
  FfnP p{};
  p.b0 = b0; p.b2 = b2; p.x = x->ptr; p.y = y->ptr; p.x_ld = x->ld; p.y_ld = y->ld;
  p.M = (int)npix(*x);
  p.tmem_cols = C <= 128 ? 128 : (C <= 256 ? 256 : 512);
  const int BNW = (C <= 256) ? C : ((C % 256 == 0) ? 256 : 128);
  FfnMaps maps;
  if (int r = encode_ffn_maps(maps, x, p.M, C, BNW, w0, w2, "ffn")) return r;
  cudaStream_t s = (cudaStream_t)stream;
  switch (C) {
    case 128: return launch_ffn<128>(maps, p, s);
    case 192: return launch_ffn<192>(maps, p, s);
    case 256: return launch_ffn<256>(maps, p, s);
    case 384: return launch_ffn<384>(maps, p, s);
    default: return launch_ffn<512>(maps, p, s);
  }
}

// Fused two-layer point-wise block without LayerNorm: y = res + W2 . ReLU(W0 . x + b0) + b2 — the 1x1 branch of ConvBlock2D
// (attentions.py:273-281) summed with the 3x3 branch's output `res`.  0 = launched, 1 = not applicable, < 0 = error.
extern "C" int s2m2b200_mlp2(const S2m2Tensor* x, const S2m2Tensor* res, const S2m2Tensor* y, const void* w0, const float* b0,
                             const void* w2, const float* b2, int32_t w_dtype, int32_t layernorm, int32_t act, void* stream) {
  S2M2_TRACE("mlp2", x, res, y, w0, b0, w2, b2, w_dtype, layernorm, act);
  return mlp2_dispatch(x, res, y, w0, b0, w2, b2, w_dtype, layernorm, act, (cudaStream_t)stream);
}
