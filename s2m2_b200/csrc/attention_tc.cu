// tcgen05 whole-row attention for the 1-D (scan-line) blocks: every key of a sequence fits on chip (N <= 384),
// so the softmax is exact in one pass over TMEM — no online rescaling.
//
//   one CTA = 128 queries of one (sequence, head):
//     phase 1  S[128 x N]   = Q K^T      tcgen05.mma (M=128, N=64 per key block, K=d), S in TMEM columns [0, N_pad)
//     phase 2  P = softmax(S)             4 warps, thread = row: TMEM -> registers -> fp16 P written to shared memory in the
//                                         K-major 128B-swizzled layout the tensor core reads as the A operand
//     phase 3  O[128 x d]   = P V         tcgen05.mma (M=128, N=d, K=64 per key block); V is consumed as it lies in memory
//                                         ([key][channel] = MN-major B operand, b_major = 1): no transpose pass; O re-uses S's
//                                         TMEM columns
//     phase 4  out = O / rowsum           TMEM -> registers -> global (full-sector stores)
//   warp 0 = TMA producer (Q once, then K blocks, then V^T blocks through one smem ring), warp 1 = MMA issuer + TMEM
//   allocator, warps 2..5 = softmax / epilogue.
// Reference op: F.scaled_dot_product_attention in the 1-D blocks (attentions.py:50,86,91,119-161).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "attention.cuh"

namespace s2m2 {

constexpr int ATC_THREADS = 192;
constexpr unsigned long long ATC_TIMEOUT = 4000000000ull;

struct AtcP {
  void* out;
  long long out_ld;
  int S, N, heads, C, kv_shift, N_pad, NKB, tmem_cols;
  float sl2;  // softmax scale * log2(e)
};
struct AtcMaps {
  CUtensorMap q;   // qkv (tokens x 3C): box {64 ch, 128 tokens}
  CUtensorMap k;   // qkv (tokens x 3C): box {64 ch, 64 tokens}
};

// ---- PTX wrappers (same idioms as conv_tc.cu) ---------------------------------------------------
namespace atc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > ATC_TIMEOUT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, SWIZZLE_128B: rows = K (keys) of 128 B = 64 MN elements; 8-row atoms SBO = 1024 B apart along K; the next
// 64-element MN chunk lies LBO = 8192 B further (one [64 keys][128 B] TMA box).  (cute: Layout_MN_SW128_Atom)
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
}  // namespace atc
using namespace atc;

template <int D>
__global__ void __launch_bounds__(ATC_THREADS, 1) attention_tc_kernel(const __grid_constant__ AtcMaps maps, const AtcP p) {
  constexpr int KC = D / 64;                       // 64-channel chunks of the head dim
  constexpr uint32_t Q_BYTES = 128 * D * 2;        // KC chunks of [128 rows][128 B]
  constexpr uint32_t BLK_BYTES = 64 * D * 2;       // one ring slot: K block or V block, both KC x [64 keys][128 B]
  constexpr int NST = (D >= 256) ? 3 : 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t r0_bytes = max(Q_BYTES, (uint32_t)(128 * p.N_pad * 2));
  uint8_t* sQP = smem;                             // Q during phase 1, P afterwards
  uint8_t* sRing = smem + ((r0_bytes + 1023) & ~1023u);
  uint64_t* bars = (uint64_t*)(sRing + NST * BLK_BYTES);   // qfull, full[NST], empty[NST], sfull, pready, ofull
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NST + 4);
  const uint32_t qfull = smem_u32(bars), full0 = smem_u32(bars + 1), empty0 = smem_u32(bars + 1 + NST);
  const uint32_t sfull = smem_u32(bars + 1 + 2 * NST), pready = sfull + 8, ofull = sfull + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, s = blockIdx.z;
  const int s_kv = (s + p.kv_shift) % p.S;
  const int NKB = p.NKB;

  if (threadIdx.x == 0) {
    mbar_init(qfull, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(sfull, 1);
    mbar_init(pready, 128);
    mbar_init(ofull, 1);
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
      mbar_expect_tx(qfull, Q_BYTES);
#pragma unroll
      for (int kc = 0; kc < KC; ++kc)
        tma_load_2d(&maps.q, smem_u32(sQP + kc * 16384), qfull, head * D + kc * 64, s * p.N + q0);
      for (int it = 0; it < 2 * NKB; ++it) {
        const int st = it % NST;
        const uint32_t ph = (it / NST) & 1;
        mbar_wait(empty0 + 8 * st, ph ^ 1);
        mbar_expect_tx(full0 + 8 * st, BLK_BYTES);
        const uint32_t dst = smem_u32(sRing + st * BLK_BYTES);
        if (it < NKB) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
            tma_load_2d(&maps.k, dst + kc * 8192, full0 + 8 * st, p.C + head * D + kc * 64, s_kv * p.N + it * 64);
        } else {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
            tma_load_2d(&maps.k, dst + kc * 8192, full0 + 8 * st, 2 * p.C + head * D + kc * 64, s_kv * p.N + (it - NKB) * 64);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // O = P V: B = V block as loaded, [64 keys][64 ch] x KC chunks = MN-major (channels contiguous): b_major bit 16
      const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(qfull, 0);
      for (int it = 0; it < NKB; ++it) {            // S[:, 64*it ..] = Q K_blk^T
        const int st = it % NST;
        mbar_wait(full0 + 8 * st, (it / NST) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t ad = make_sdesc(smem_u32(sQP + kc * 16384)), bd = make_sdesc(smem_u32(sRing + st * BLK_BYTES + kc * 8192));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + it * 64, ad + 2 * k, bd + 2 * k, idesc_s, (kc | k) != 0);
        }
        umma_commit(empty0 + 8 * st);
      }
      umma_commit(sfull);
      mbar_wait(pready, 0);                          // P is in shared memory, S has been read out of TMEM
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int b = 0; b < NKB; ++b) {                // O += P[:, 64 b ..] V_blk
        const int it = NKB + b, st = it % NST;
        mbar_wait(full0 + 8 * st, (it / NST) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t ad = make_sdesc(smem_u32(sQP + b * 16384)), bd = make_sdesc_mn(smem_u32(sRing + st * BLK_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)   // A: +32 B per 16 keys within the 128 B row; B: +16 key rows = 2 swizzle atoms = 2048 B
          umma_f16(tmem_base, ad + 2 * k, bd + (2048 >> 4) * k, idesc_o, (b | k) != 0);
        umma_commit(empty0 + 8 * st);
      }
      umma_commit(ofull);
    }
  } else {
    // ================= softmax + epilogue: thread = one query row =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int t = q0 + r;
    const bool valid = t < p.N;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(sfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // pass 1: row max over the valid keys (32 columns per TMEM load: fewer load/wait round trips)
    float m = -3.0e38f;
    for (int c0 = 0; c0 < p.N_pad; c0 += 32) {
      uint32_t rr[32];
      tmem_ld32(trow + c0, rr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < p.N) m = fmaxf(m, __uint_as_float(rr[i]));
    }
    // pass 2: p = exp2((s - m) * scale*log2e), row sum, fp16 P into the swizzled A-operand layout
    float l = 0.f;
    const float msc = m * p.sl2;
    for (int c0 = 0; c0 < p.N_pad; c0 += 32) {
      uint32_t rr[32];
      tmem_ld32(trow + c0, rr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __half2 h[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float a = (c0 + i < p.N) ? exp2f(fmaf(__uint_as_float(rr[i]), p.sl2, -msc)) : 0.f;
        const float b = (c0 + i + 1 < p.N) ? exp2f(fmaf(__uint_as_float(rr[i + 1]), p.sl2, -msc)) : 0.f;
        l += a + b;
        h[i >> 1] = __floats2half2_rn(a, b);
      }
      uint8_t* blk = sQP + (c0 >> 6) * 16384 + r * 128;            // key block, this row
      const int u0 = (c0 & 63) >> 3;                                  // first 16-byte unit of this chunk within the 128 B row
#pragma unroll
      for (int u = 0; u < 4; ++u) *(uint4*)(blk + (((u0 + u) ^ (r & 7)) << 4)) = *(uint4*)&h[4 * u];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy P writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // our TMEM reads are ordered before the PV MMAs
    mbar_arrive(pready);

    // epilogue: O / l -> fp16 -> global, full 32 B sectors per store instruction via lane-pair exchange
    mbar_wait(ofull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.f / l;
    const bool pvalid = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
    const bool odd = lane & 1;
    __half* ob = (__half*)p.out;
    const long long off0 = ((long long)s * p.N + t) * p.out_ld + head * D;
    const long long poff0 = __shfl_xor_sync(0xffffffffu, off0, 1);
#pragma unroll 1
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint32_t rr[16];
      tmem_ld16(trow + c0, rr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __half2 h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(__uint_as_float(rr[2 * i]) * inv, __uint_as_float(rr[2 * i + 1]) * inv);
      const uint4 lo = *(uint4*)&h[0], hi = *(uint4*)&h[4];
      const uint4 send = odd ? lo : hi;
      uint4 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
      recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
      if (odd ? pvalid : valid) *(uint4*)(ob + (odd ? poff0 + c0 + 8 : off0 + c0)) = odd ? recv : lo;
      if (odd ? valid : pvalid) *(uint4*)(ob + (odd ? off0 + c0 + 8 : poff0 + c0)) = odd ? hi : recv;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// ---- host ----------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 atc_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)f;
  }
  return fn;
}

int attention_tc_supported(const AttnP& p) {
  if (!p.qh || !p.oh || p.pesum) return 0;
  if (!(p.d == 64 || p.d == 128 || p.d == 256)) return 0;
  if (p.N < 48 || p.N > 384) return 0;
  if (p.qkv_ld % 8 || p.out_ld % 8 || !aligned(p.qkv, 16) || !aligned(p.out, 16)) return 0;
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    ok = (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
          major == 10 && atc_encode() != nullptr) ? 1 : 0;
  }
  return ok;
}

size_t attention_tc_workspace_bytes(const AttnP& p) { (void)p; return 16; }   // no scratch needed any more (kept for the ABI)

template <int D>
static int launch_atc(const AtcMaps& maps, const AtcP& p, dim3 grid, size_t smem, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    S2M2_CHECK(e == cudaSuccess, "attention_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  launch_k(attention_tc_kernel<D>, grid, ATC_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("attention_tc");
  return 0;
}

int attention_tc_launch(const AttnP& a, void* workspace, cudaStream_t s) {
  AtcP p{};
  const int D = a.d;
  p.out = a.out; p.out_ld = a.out_ld; p.S = a.S; p.N = a.N; p.heads = a.heads; p.C = a.C; p.kv_shift = a.kv_shift;
  p.N_pad = (a.N + 63) / 64 * 64; p.NKB = p.N_pad / 64;
  const int need = p.N_pad > D ? p.N_pad : D;
  p.tmem_cols = need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512));
  p.sl2 = a.scale * 1.4426950408889634f;
  (void)workspace;
  AtcMaps maps;
  const long long tokens = (long long)a.S * a.N;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(3 * a.C), (cuuint64_t)tokens}, strides[1] = {(cuuint64_t)a.qkv_ld * 2};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t boxq[2] = {64, 128}, boxk[2] = {64, 64};
    CUresult r1 = atc_encode()(&maps.q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.qkv), dims, strides, boxq, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = atc_encode()(&maps.k, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.qkv), dims, strides, boxk, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, "attention_tc: cuTensorMapEncodeTiled(qkv) failed: %d %d", (int)r1, (int)r2);
  }
  const int NST = D >= 256 ? 3 : 4;
  const size_t r0 = (size_t)128 * 2 * (p.N_pad > D ? p.N_pad : D);
  const size_t smem = ((r0 + 1023) & ~(size_t)1023) + (size_t)NST * 64 * D * 2 + (2 * NST + 4) * 8 + 16 + 1024;
  S2M2_CHECK(smem <= 220 * 1024, "attention_tc: shared memory %zu too large", smem);
  dim3 grid((a.N + 127) / 128, a.heads, a.S);
  if (D == 256) return launch_atc<256>(maps, p, grid, smem, s);
  if (D == 128) return launch_atc<128>(maps, p, grid, smem, s);
  return launch_atc<64>(maps, p, grid, smem, s);
}

}  // namespace s2m2
