// tcgen05 implicit-GEMM convolution for sm_100a: fp16 operands, fp32 accumulation in TMEM.
//
//   D[128 pixels x BN couts] += A[128 x 64] * B[BN x 64]^T   per k-block (64 input channels of one filter tap)
//
//   * A (activations, NHWC): one TMA box {64 ch, tw, th, 1 img} of the input at the tap's spatial offset — TMA's
//     out-of-bounds zero fill IS the convolution's zero padding, so there is no im2col and no border code.
//     The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle = the canonical K-major
//     UMMA layout.  conv over cat([x, x2]) walks two tensor maps (no concat buffer).
//   * B (weights, [Cout][taps*Cin] K-major): TMA box {64, BN}.
//   * warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue
//     (tcgen05.ld -> bias/activation/residual -> 16/32-byte global stores, optional ConvTranspose2x2 scatter).
//   * smem ring of STAGES k-blocks guarded by full/empty mbarriers; tcgen05.commit releases slots.
//
// Reference ops replaced: the big nn.Conv2d / nn.Linear / ConvTranspose2d(k=2) layers (include/s2m2b200.h).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"

namespace s2m2 {

constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 192;
constexpr unsigned long long TC_TIMEOUT_CYCLES = 4000000000ull;  // ~2 s: trap instead of hanging the GPU

struct TcP {
  const float* bias;
  const void* res;
  void* y;
  long long res_ld, y_ld;
  int rh, yh;            // residual / output fp16 flags
  int flat;              // 1x1 conv: pixels are one flat dimension
  int Ho, Wo, Cout;      // output geometry (for flat: Wo = total pixels, Ho = 1)
  int tw, th, tiles_x, tiles_y;
  int KW, ph, pw, taps, stride;
  int cpb1, cpb2;        // 64-channel chunks per tap from x and x2
  int act, shuffle, Cr;
  int tma_res;           // (with tma_store) the residual tile is TMA-loaded into the staging buffer instead of scattered global loads
  int tma_store;         // epilogue stages the tile in shared memory and writes it with cp.async.bulk.tensor (fp16, no scatter)
  int dbg;               // S2M2_TC_DBG experiments: 1 = no global stores, 2 = epilogue only drains (no math / stores)
  const void* res2;      // epilogue modes 2 / 3: second tile operand (fp16), read with plain loads
  long long res2_ld;
  int emode, res_from;   // include/s2m2b200.h: 0 add, 1 multiply (columns >= res_from), 2 gate blend, 3 GRU update
  int m_major;           // persistent kernel: tile order (see tile_coords)
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
    if (clock64() - t0 > TC_TIMEOUT_CYCLES) __trap();  // surfaces as a launch failure instead of a hung GPU
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// ---- CTA-pair (cta_group::2) variants: the pair's TMA loads complete on the LEADER's mbarrier (cluster address with the peer
// bit cleared), one tcgen05.mma drives both SMs, its commit is multicast to the same barrier offset in both CTAs.
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {   // shared::cluster address of the same offset in CTA rank 0
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
  return remote;
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's copy of a barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// waits for this thread's outstanding tcgen05.ld's; the registers pass through the asm so that no use of them can be scheduled
// above the wait (tcgen05.ld is asynchronous: its destination registers are undefined until wait::ld)
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Tensor maps of one launch: act[0..3] = activation sources (stride 1: x, x2; stride 2: the four (row, column)
// parity sub-sampled views of x, so that a strided convolution is a walk over plain strided views), w = weights.
struct TcMaps {
  CUtensorMap act[4];
  CUtensorMap w;
  CUtensorMap out;   // v2 epilogue: box {BN/4 channels, 32 pixels of the tile} for the TMA store
  CUtensorMap res;   // same box over the residual tensor
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcP p) {
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2, B_BYTES = BN * TC_BK * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = (uint64_t*)(sB + STAGES * B_BYTES);  // full[STAGES], empty[STAGES], tmem_full
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);

  // ---- tile coordinates
  const int n0 = blockIdx.y * BN;
  int img = 0, x0 = 0, y0 = 0;
  if (p.flat) {
    x0 = blockIdx.x * TC_BM;
  } else {
    int mt = blockIdx.x;
    int txi = mt % p.tiles_x;
    int t = mt / p.tiles_x;
    int tyi = t % p.tiles_y;
    img = t / p.tiles_y;
    x0 = txi * p.tw;
    y0 = tyi * p.th;
  }
  const int cpb = p.cpb1 + p.cpb2;
  const int nkb = p.taps * cpb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation: BN fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN));
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
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, A_BYTES + B_BYTES);
        const int tap = kb / cpb, ch = kb - tap * cpb;
        const int ky = tap / p.KW, kx = tap - ky * p.KW;
        const uint32_t dstA = smem_u32(sA + s * A_BYTES), dstB = smem_u32(sB + s * B_BYTES);
        const int dx = kx - p.pw, dy = ky - p.ph;
        if (p.stride == 1) {
          if (ch < p.cpb1) tma_load_4d(&maps.act[0], dstA, full0 + 8 * s, ch * TC_BK, x0 + dx, y0 + dy, img);
          else tma_load_4d(&maps.act[1], dstA, full0 + 8 * s, (ch - p.cpb1) * TC_BK, x0 + dx, y0 + dy, img);
        } else {  // stride 2: input pixel (2y+dy, 2x+dx) lives in parity view (dy&1, dx&1) at (y + floor(dy/2), x + floor(dx/2))
          tma_load_4d(&maps.act[(dy & 1) * 2 + (dx & 1)], dstA, full0 + 8 * s, ch * TC_BK, x0 + (dx >> 1), y0 + (dy >> 1), img);
        }
        tma_load_2d(&maps.w, dstB, full0 + 8 * s, kb * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t ad = make_sdesc(smem_u32(sA + s * A_BYTES)), bd = make_sdesc(smem_u32(sB + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)  // +32 B (= 2 in the >>4 encoded start address) per 16-element K step
          umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        umma_commit(empty0 + 8 * s);            // frees the smem slot when these MMAs retire
      }
      umma_commit(tfull);                        // accumulator complete
    }
  } else {
    // ================= epilogue warps (TMEM lanes 32*(warp%4) ..) =================
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // accumulator row = pixel within the tile
    bool valid;
    long long opix;                              // output pixel index (before shuffle)
    int py = 0, px = 0;
    if (p.flat) {
      valid = x0 + r < p.Wo;
      opix = x0 + r;
    } else {
      py = y0 + r / p.tw;
      px = x0 + r % p.tw;
      valid = (r < p.tw * p.th) && py < p.Ho && px < p.Wo;
      opix = ((long long)img * p.Ho + py) * p.Wo + px;
    }
    mbar_wait(tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      const int n = n0 + c0;
      if (n >= p.Cout) break;                    // uniform across the warp
      uint32_t rr[16];
      tmem_ld16(trow + c0, rr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!valid) continue;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float t = __uint_as_float(rr[i]);
        if (p.bias) t += __ldg(p.bias + n + i);
        v[i] = act_apply(t, p.act);
      }
      long long off;
      if (p.shuffle) {  // ConvTranspose2d(k=2,s=2) scatter: column = (a*2+b)*Cr + co
        const int qd = n / p.Cr, co = n - qd * p.Cr;
        const long long pix = ((long long)img * 2 * p.Ho + 2 * py + (qd >> 1)) * (2 * p.Wo) + 2 * px + (qd & 1);
        off = pix * p.y_ld + co;
      } else {
        off = opix * p.y_ld + n;
        if (p.res) {
          const long long ro = opix * p.res_ld + n;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 t = ld4(p.res, ro + i, p.rh);
            v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
          }
        }
      }
      if (p.yh) {
        __half2 h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        uint4* dst = (uint4*)((__half*)p.y + off);
        dst[0] = *(uint4*)&h[0];
        dst[1] = *(uint4*)&h[4];
      } else {
        float4* dst = (float4*)((float*)p.y + off);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN));
  }
}


// =====================================================================================================
// v2: persistent kernel.  One CTA per SM loops over output tiles (n-tile major so concurrently running CTAs
// share the weight tile in L2); two TMEM accumulators so the epilogue of tile i overlaps the mainloop of tile
// i+1; 16 epilogue warps (4 per TMEM lane quarter, each a quarter of the columns — the GELU epilogue, not the
// MMA, is the critical path of the K=256 GEMMs); the smem ring keeps streaming across tile boundaries.
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17 = epilogue.
// =====================================================================================================
constexpr int TC2_EPI_GROUPS = 4;                       // epilogue warps per TMEM lane quarter
constexpr int TC2_THREADS = 64 + 128 * TC2_EPI_GROUPS;   // 576

// PAIR: two CTAs of a cluster (one TPC) work on two vertically adjacent 128-pixel tiles with the same output channels as ONE
// 256 x BN tcgen05.mma.cta_group::2: each CTA loads its own A tile but only half of the weight tile, so the L2->SMEM operand
// traffic per FLOP (the limiter of the 3x3 convolutions) drops by a third and the ring gets deeper.
// BRES (weights resident): layers whose whole weight matrix is one narrow tile of a few k-blocks (64-wide 3x3 / 1x1 layers at full
// and half resolution: K <= 576, one n-tile) load it ONCE per CTA into its own shared-memory region instead of streaming it
// through the ring with every pixel tile — every CTA re-reading the same 8-72 KB from the same few L2 slices was what held these
// layers at 55 GB/s per SM (profiles/r2_kernels_ncu2.md: the epilogue warps wait for the accumulator, the ring for its loads).
constexpr int TC2_BRES_MAX_KB = 9;
template <int BN, int STAGES, bool PAIR, bool BRES = false>
__global__ void __launch_bounds__(TC2_THREADS, 1) conv_tc_persistent_kernel(const __grid_constant__ TcMaps maps, const TcP p,
                                                                            int m_tiles, int n_tiles) {
  static_assert(!(PAIR && BRES), "weights-resident variant is single-CTA");
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2, B_BYTES = (PAIR ? BN / 2 : BN) * TC_BK * 2;
  constexpr int B_SLOTS = BRES ? TC2_BRES_MAX_KB : STAGES;
  uint32_t rank = 0;
  if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  constexpr uint32_t STG_BYTES = 32 * (BN / TC2_EPI_GROUPS) * 2;          // per epilogue warp: 32 rows x CG fp16
  uint8_t* sStage = sB + B_SLOTS * B_BYTES;                                 // 16 warps x STG_BYTES, 1024-aligned
  uint64_t* bars = (uint64_t*)(sStage + 4 * TC2_EPI_GROUPS * 1024 * ((STG_BYTES + 1023) / 1024));  // full[S], empty[S], tfull[2], tempty[2], rbar[16], bres
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4 + 4 * TC2_EPI_GROUPS + 1);
  const uint32_t rbar0 = smem_u32(bars + 2 * STAGES + 4);
  const uint32_t bres_bar = smem_u32(bars + 2 * STAGES + 4 + 4 * TC2_EPI_GROUPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);
  const int cpb = p.cpb1 + p.cpb2;
  const int nkb = p.taps * cpb;
  // work items: tiles, or (PAIR) pairs of m-tiles (2i, 2i+1) x one n-tile, dealt to CTAs / clusters round-robin
  const int m_items = PAIR ? (m_tiles + 1) / 2 : m_tiles;
  const int total_tiles = m_items * n_tiles;
  const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4 * TC2_EPI_GROUPS * (PAIR ? 2 : 1));   // one arrive per epilogue warp (of both CTAs)
    }
    for (int a = 0; a < 4 * TC2_EPI_GROUPS; ++a) mbar_init(rbar0 + 8 * a, 1);
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  // tile id -> coordinates.  m-major (default): tile = mt * n_tiles + nt — the n-tiles of one pixel tile run at the same time on
  // neighbouring CTAs, so the activations are fetched from HBM once and re-read from L2 (the whole weight matrix is a few MB and
  // L2-resident either way).  n-major (tile = nt * m_tiles + mt, round 1) re-streams the activations from HBM once per n-tile as soon
  // as they exceed L2 — 3x for the 768-wide gate+fusion layers at batch 8.
  auto tile_coords = [&](int tile, int& n0, int& img, int& x0, int& y0) -> bool {
    int nt, mt;
    if (p.m_major) { mt = tile / n_tiles; nt = tile - mt * n_tiles; }
    else { nt = tile / m_items; mt = tile - nt * m_items; }
    if constexpr (PAIR) mt = 2 * mt + (int)rank;
    n0 = nt * BN;
    img = 0; y0 = 0;
    if (p.flat) {
      x0 = mt * TC_BM;
    } else {
      const int txi = mt % p.tiles_x;
      const int t = mt / p.tiles_x;
      const int tyi = t % p.tiles_y;
      img = t / p.tiles_y;
      x0 = txi * p.tw;
      y0 = tyi * p.th;
    }
    return mt < m_tiles;   // PAIR, odd tile count: the last pair's second tile lies outside the tensor (TMA zero-fills / clips)
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;  // running k-block counter across tiles
      auto load_a = [&](const CUtensorMap* m, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
        if constexpr (PAIR) tma_load_4d_pair(m, dst, bar, c0, c1, c2, c3);
        else tma_load_4d(m, dst, bar, c0, c1, c2, c3);
      };
      if constexpr (BRES) {   // the whole weight tile, once (n_tiles == 1: it never changes)
        if (first_tile < total_tiles) {
          mbar_expect_tx(bres_bar, (uint32_t)nkb * B_BYTES);
          for (int kb = 0; kb < nkb; ++kb) tma_load_2d(&maps.w, smem_u32(sB + kb * B_BYTES), bres_bar, kb * TC_BK, 0);
        }
      }
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        int n0, img, x0, y0;
        tile_coords(tile, n0, img, x0, y0);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);                       // own slot free (the commit is multicast to both CTAs)
          if constexpr (PAIR) {
            if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2 * (A_BYTES + B_BYTES));   // leader's barrier counts both CTAs' bytes
          } else {
            mbar_expect_tx(full0 + 8 * s, BRES ? A_BYTES : A_BYTES + B_BYTES);
          }
          const int tap = kb / cpb, ch = kb - tap * cpb;
          const int ky = tap / p.KW, kx = tap - ky * p.KW;
          const uint32_t dstA = smem_u32(sA + s * A_BYTES), dstB = smem_u32(sB + (BRES ? 0 : s) * B_BYTES);
          const int dx = kx - p.pw, dy = ky - p.ph;
          if (p.stride == 1) {
            if (ch < p.cpb1) load_a(&maps.act[0], dstA, full0 + 8 * s, ch * TC_BK, x0 + dx, y0 + dy, img);
            else load_a(&maps.act[1], dstA, full0 + 8 * s, (ch - p.cpb1) * TC_BK, x0 + dx, y0 + dy, img);
          } else {
            load_a(&maps.act[(dy & 1) * 2 + (dx & 1)], dstA, full0 + 8 * s, ch * TC_BK, x0 + (dx >> 1), y0 + (dy >> 1), img);
          }
          if constexpr (PAIR) tma_load_2d_pair(&maps.w, dstB, full0 + 8 * s, kb * TC_BK, n0 + (int)rank * (BN / 2));   // this CTA's half of the weight rows
          else if constexpr (!BRES) tma_load_2d(&maps.w, dstB, full0 + 8 * s, kb * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && rank == 0) {   // PAIR: the leader issues for both SMs
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * TC_BM : TC_BM) >> 4) << 24);
      uint32_t it = 0, lt = 0;  // k-block counter, local tile counter
      if constexpr (BRES) {
        if (first_tile < total_tiles) mbar_wait(bres_bar, 0);
      }
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++lt) {
        const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);          // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ad = make_sdesc(smem_u32(sA + s * A_BYTES)), bd = make_sdesc(smem_u32(sB + (BRES ? kb : s) * B_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            if constexpr (PAIR) umma_f16_pair(tacc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
            else umma_f16(tacc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (PAIR) umma_commit_pair(empty0 + 8 * s);
          else umma_commit(empty0 + 8 * s);
        }
        if constexpr (PAIR) umma_commit_pair(tfull0 + 8 * acc);
        else umma_commit(tfull0 + 8 * acc);
      }
    }
  } else {
    // ================= epilogue: warp%4 = TMEM lane quarter, (warp-2)/4 = column group =================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int CG = BN / TC2_EPI_GROUPS;   // columns per group
    const int r = q * 32 + lane;
    // TMA-store staging: this warp's 32 rows x CG columns, rows of CG*2 bytes, swizzled like the store tensor map
    constexpr uint32_t STG_STRIDE = 1024 * ((32 * CG * 2 + 1023) / 1024);
    constexpr uint32_t SWZ_MASK = (CG * 2 >= 128) ? 7u : ((CG * 2 == 64) ? 3u : 1u);
    uint8_t* stg = sStage + (warp - 2) * STG_STRIDE;
    const uint32_t stg_u32 = smem_u32(stg);
    uint32_t lt = 0, rt = 0;   // tiles processed, residual tiles loaded (phase of this warp's residual barrier)
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++lt) {
      int n0, img, x0, y0;
      const bool tile_ok = tile_coords(tile, n0, img, x0, y0);
      const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
      bool valid;
      long long opix;
      int py = 0, px = 0;
      if (!tile_ok) {
        valid = false; opix = 0;
      } else if (p.flat) {
        valid = x0 + r < p.Wo;
        opix = x0 + r;
      } else {
        py = y0 + r / p.tw;
        px = x0 + r % p.tw;
        valid = py < p.Ho && px < p.Wo;
        opix = ((long long)img * p.Ho + py) * p.Wo + px;
      }
      // the residual does not depend on the accumulator: fetch this warp's slice before waiting for the MMAs
      constexpr int NCH = CG / 16;                       // 16-column chunks per warp
      uint4 rpre[NCH][2];
      const bool has_res = p.res != nullptr && !p.shuffle;
      if (has_res && p.rh && !p.tma_res && p.emode == 0) {
        // full-sector reads, same lane-pair trick as the stores: instruction 1 reads the even row of the pair (low half by
        // the even lane, high half by the odd lane), instruction 2 the odd row; one exchange puts the halves back in place.
        const bool odd = lane & 1;
        const bool pv = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
        const long long popix = __shfl_xor_sync(0xffffffffu, opix, 1);
        const long long row_e = odd ? popix : opix, row_o = odd ? opix : popix;   // even / odd row of the pair
        const bool ve = odd ? pv : valid, vo = odd ? valid : pv;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int n = n0 + half * CG + c * 16;
          uint4 a = make_uint4(0, 0, 0, 0), b = a;
          if (n < p.Cout) {
            const __half* rb = (const __half*)p.res + n + (odd ? 8 : 0);
            if (ve) a = *(const uint4*)(rb + row_e * p.res_ld);   // plain loads: the output may alias the residual (z += f(z))
            if (vo) b = *(const uint4*)(rb + row_o * p.res_ld);
          }
          // even lane holds (own.lo, partner.lo), odd lane holds (partner.hi, own.hi): swap the foreign halves
          const uint4 send = odd ? a : b;
          uint4 recv;
          recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
          recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
          rpre[c][0] = odd ? recv : a;     // own low half
          rpre[c][1] = odd ? b : recv;     // own high half
        }
      }
      if (p.tma_store) {   // the previous tile's bulk store must have read the staging tile before it is overwritten
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (p.tma_res && tile_ok && n0 + half * CG < p.Cout && n0 + half * CG >= p.res_from) {   // residual tile (same box as the store) -> staging buffer
            const int r0 = q * 32;
            int cx, cy;
            if (p.flat) { cx = x0 + r0; cy = 0; }
            else { cx = x0 + r0 % p.tw; cy = y0 + r0 / p.tw; }
            mbar_expect_tx(rbar0 + 8 * (warp - 2), 32 * CG * 2);
            tma_load_4d(&maps.res, stg_u32, rbar0 + 8 * (warp - 2), n0 + half * CG - p.res_from, cx, cy, img);
          }
        }
        __syncwarp();
      }
      mbar_wait(tfull0 + 8 * acc, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const bool use_tma_res = p.tma_res && tile_ok && n0 + half * CG >= p.res_from;
      if (use_tma_res && n0 + half * CG < p.Cout) {
        mbar_wait(rbar0 + 8 * (warp - 2), rt & 1);
        ++rt;
      }
      if (p.dbg & 2) valid = false;
      // neighbour-lane exchange for full-sector stores (see below): lane l pairs with lane l^1
      const bool pvalid = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
      // the TMEM loads are double-buffered: chunk c+1 is in flight while chunk c goes through the activation / stores
      uint32_t rbuf[2][16];
      if (n0 + half * CG < p.Cout) tmem_ld16(trow + half * CG, rbuf[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int c0 = half * CG + c * 16;
        const int n = n0 + c0;
        if (n >= p.Cout) break;
        uint32_t* rr = rbuf[c & 1];
        tmem_ld_wait16(rr);
        if (c + 1 < NCH && n + 16 < p.Cout) tmem_ld16(trow + c0 + 16, rbuf[(c + 1) & 1]);
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
        // the activation switch is hoisted out of the element loop: each case is a straight-line, fully unrolled block
        // (16 independent dependency chains) instead of 16 branchy evaluations
        switch (p.act) {
          case S2M2_ACT_GELU:
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);
            break;
          case S2M2_ACT_RELU:
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
            break;
          case S2M2_ACT_SIGMOID:
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fdividef(1.0f, 1.0f + __expf(-v[i]));
            break;
          case S2M2_ACT_TANH:
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v[i]));   // tanh, ~1e-6 abs
            break;
          default: break;
        }
        long long off;
        if (p.shuffle) {
          const int qd = n / p.Cr, co = n - qd * p.Cr;
          const long long pix = ((long long)img * 2 * p.Ho + 2 * py + (qd >> 1)) * (2 * p.Wo) + 2 * px + (qd & 1);
          off = pix * p.y_ld + co;
        } else {
          off = opix * p.y_ld + n;
          if (use_tma_res) {
            // emode 0: + R1;  1: * R1;  2: gate blend g R1 + (1 - g) R2;  3: GRU update R1 q + (1 - R1) R2   (s2m2b200.h)
            uint4 r2v[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if (p.emode >= 2 && valid) {
              const uint4* r2p = (const uint4*)((const __half*)p.res2 + opix * p.res2_ld + n);
              r2v[0] = r2p[0]; r2v[1] = r2p[1];                       // plain loads: y may alias R2 (in-place state update)
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const uint32_t lin = (uint32_t)lane * (CG * 2) + (uint32_t)(c * 2 + u) * 16;
              const uint4 rv = *(const uint4*)(stg + (lin ^ (((lin >> 7) & SWZ_MASK) << 4)));
              const __half2* hr = (const __half2*)&rv;
              const __half2* h2 = (const __half2*)&r2v[u];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hr[i]);
                float& a0 = v[8 * u + 2 * i];
                float& a1 = v[8 * u + 2 * i + 1];
                if (p.emode == 0) { a0 += f.x; a1 += f.y; }
                else if (p.emode == 1) { a0 *= f.x; a1 *= f.y; }
                else {
                  const float2 g = __half22float2(h2[i]);
                  if (p.emode == 2) {
                    const float w0 = fminf(fmaxf(a0, 0.01f), 0.99f), w1 = fminf(fmaxf(a1, 0.01f), 0.99f);
                    a0 = w0 * f.x + (1.0f - w0) * g.x; a1 = w1 * f.y + (1.0f - w1) * g.y;
                  } else {
                    a0 = f.x * a0 + (1.0f - f.x) * g.x; a1 = f.y * a1 + (1.0f - f.y) * g.y;
                  }
                }
              }
            }
          } else if (has_res && valid && p.emode == 0) {
            if (p.rh) {
              const __half2* hr = (const __half2*)&rpre[c][0];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 f = __half22float2(hr[i]);
                v[2 * i] += f.x; v[2 * i + 1] += f.y;
              }
            } else {
              const long long ro = opix * p.res_ld + n;
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float4 t = ld4(p.res, ro + i, 0);
                v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
              }
            }
          }
        }
        if (p.dbg & 1) {
          if (v[0] == 123.456f) ((float*)p.y)[0] = v[1];   // keep the math alive without storing
        } else if (p.tma_store) {
          // stage this thread's 16 columns (two 16-byte units of its row) in the swizzled staging tile
          __half2 h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t lin = (uint32_t)lane * (CG * 2) + (uint32_t)(c * 2 + u) * 16;
            *(uint4*)(stg + (lin ^ (((lin >> 7) & SWZ_MASK) << 4))) = *(uint4*)&h[4 * u];
          }
        } else if (p.yh) {
          // fp16: a thread's 16 columns are one 32-byte sector = two 16-byte halves.  Writing them as two scattered
          // 16-byte stores makes every store a partial-sector write; instead lanes l and l^1 swap one half so that
          // each store instruction writes both halves of a sector from adjacent lanes (full 32 B sectors).
          __half2 h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
          const uint4 lo = *(uint4*)&h[0], hi = *(uint4*)&h[4];
          const bool odd = lane & 1;
          const uint4 send = odd ? lo : hi;
          uint4 recv;
          recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
          recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
          const long long poff = __shfl_xor_sync(0xffffffffu, off, 1);     // partner row's element offset
          __half* yb = (__half*)p.y;
          // store 1 -> even row of the pair: even lane writes its own low half, odd lane the even row's high half
          if (odd ? pvalid : valid) *(uint4*)(yb + (odd ? poff + 8 : off)) = odd ? recv : lo;
          // store 2 -> odd row of the pair: even lane writes the odd row's low half, odd lane its own high half
          if (odd ? valid : pvalid) *(uint4*)(yb + (odd ? off + 8 : poff)) = odd ? hi : recv;
        } else if (valid) {
          float4* dst = (float4*)((float*)p.y + off);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
      // all of this warp's TMEM reads have completed (wait::ld above): hand the accumulator back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (p.tma_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_leader(tempty0 + 8 * acc);   // the leader's MMA thread waits for both CTAs' epilogues
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * acc) : "memory");
        if (p.tma_store && tile_ok && n0 + half * CG < p.Cout && !(p.dbg & 3)) {
          // one bulk tensor store: 32 pixels (rows q*32.. of the tile) x CG channels; TMA clips everything outside the tensor
          const int r0 = q * 32;
          int cx, cy;
          if (p.flat) { cx = x0 + r0; cy = 0; }
          else { cx = x0 + r0 % p.tw; cy = y0 + r0 / p.tw; }
          tma_store_4d(&maps.out, stg_u32, n0 + half * CG, cx, cy, img);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) cluster_sync_all();   // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrives are in flight
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
  }
}

// ---- host side -----------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  if (plan_without_device()) return (PFN_cuTensorMapEncodeTiled_v12000)plan_encode_stub();
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

static bool tc_device_ok() {
  if (plan_without_device()) return true;
  static int ok_of[S2M2_MAX_DEVICES] = {0};   // 0 = unknown, 1 = yes, 2 = no  (per device)
  const int di = current_device_index();
  if (ok_of[di] == 0) {
    int major = 0;
    ok_of[di] = (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, di) == cudaSuccess && major == 10 && get_encode() != nullptr) ? 1 : 2;
  }
  return ok_of[di] == 1;
}

int conv_tc_supported(const S2m2Conv* a) {
  const S2m2Tensor &x = a->x, &x2 = a->x2, &y = a->y, &res = a->res;
  if (x.dtype != S2M2_F16 || a->w_dtype != S2M2_F16) return 0;
  if (a->batch > 1) return 0;
  if (a->stride == 2) {
    // odd square kernels with 'same' padding (3x3 pad 1: conv2_down; 5x5 pad 2: conv1_down) via the four parity views
    if (a->KH != a->KW || (a->KH != 3 && a->KH != 5) || a->pad_h != a->KH / 2 || a->pad_w != a->KW / 2 || x.H % 2 || x.W % 2 || x2.ptr ||
        a->shuffle2x2)
      return 0;
  } else if (a->stride != 1) {
    return 0;
  }
  if (x.C % TC_BK || x.ld % 8 || !aligned(x.ptr, 16)) return 0;
  if (x2.ptr && (x2.C % TC_BK || x2.ld % 8 || !aligned(x2.ptr, 16))) return 0;
  if (!aligned(a->w, 16) || (a->bias && !aligned(a->bias, 16))) return 0;
  if (a->Cout % 16 || a->Cout < 16) return 0;   // narrow heads still use a 64-wide tile (TMA zero-fills the missing rows)
  // same-size convolutions only (pad = k/2) or 1x1
  if (a->stride == 1) {
    const int Ho = x.H + 2 * a->pad_h - a->KH + 1, Wo = x.W + 2 * a->pad_w - a->KW + 1;
    if (Ho != x.H || Wo != x.W) return 0;
  }
  const int ye = y.dtype == S2M2_F16 ? 8 : 4;  // elements per 16 B
  if (y.ld % ye || !aligned(y.ptr, 16)) return 0;
  if (a->shuffle2x2) {
    if ((a->Cout / 4) % 16) return 0;
  } else if (res.ptr) {
    if (res.ld % 8 || !aligned(res.ptr, 16)) return 0;
  }
  if ((long long)x.N * x.H * x.W < 256) return 0;   // tiny problems: launch-bound either way
  if (a->emode != 0) {   // fused gate / GRU epilogues ride the TMA-store + TMA-residual path of the persistent kernel
    if (y.dtype != S2M2_F16 || a->shuffle2x2 || !res.ptr || res.dtype != S2M2_F16 || a->stride != 1) return 0;
    if (a->emode == 1 && (a->res_from % 64 || a->res_from < 0 || a->res_from >= a->Cout)) return 0;
    if (a->emode != 1 && a->res_from != 0) return 0;
    if (a->emode >= 2 && (!a->res2.ptr || a->res2.dtype != S2M2_F16 || a->res2.ld % 8 || !aligned(a->res2.ptr, 16))) return 0;
    if (getenv("S2M2_TC_V1") || getenv("S2M2_TC_NO_TMA_STORE") || getenv("S2M2_TC_NO_TMA_RES")) return 0;
  }
  return tc_device_ok() ? 1 : 0;
}

static int encode_act(CUtensorMap* m, const S2m2Tensor& t, bool flat, int tw, int th, int parity = -1) {
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  void* base = t.ptr;
  if (parity >= 0) {  // every second row / column starting at (parity>>1, parity&1)
    const int py = parity >> 1, px = parity & 1;
    base = (char*)t.ptr + ((long long)py * t.W + px) * t.ld * 2;
    dims[0] = t.C; dims[1] = t.W / 2; dims[2] = t.H / 2; dims[3] = t.N;
    strides[0] = (cuuint64_t)t.ld * 2 * 2; strides[1] = (cuuint64_t)t.ld * 2 * t.W * 2; strides[2] = (cuuint64_t)t.ld * 2 * t.W * t.H;
    box[0] = TC_BK; box[1] = tw; box[2] = th; box[3] = 1;
  } else if (flat) {
    dims[0] = t.C; dims[1] = (cuuint64_t)t.N * t.H * t.W; dims[2] = 1; dims[3] = 1;
    strides[0] = (cuuint64_t)t.ld * 2; strides[1] = strides[0] * dims[1]; strides[2] = strides[1];
    box[0] = TC_BK; box[1] = TC_BM; box[2] = 1; box[3] = 1;
  } else {
    dims[0] = t.C; dims[1] = t.W; dims[2] = t.H; dims[3] = t.N;
    strides[0] = (cuuint64_t)t.ld * 2; strides[1] = strides[0] * t.W; strides[2] = strides[1] * t.H;
    box[0] = TC_BK; box[1] = tw; box[2] = th; box[3] = 1;
  }
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  S2M2_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return 0;
}

template <int BN, int STAGES>
static int launch_tc(const TcMaps& maps, const TcP& p, dim3 grid, cudaStream_t s) {
  constexpr size_t smem = STAGES * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool attr[S2M2_MAX_DEVICES] = {false};
  const int di = current_device_index();
  if (!attr[di]) {
    cudaError_t e = set_func_attr(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    S2M2_CHECK(e == cudaSuccess, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr[di] = true;
  }
  launch_k(conv_tc_kernel<BN, STAGES>, grid, TC_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("conv_tc");
  return 0;
}

template <int BN, int STAGES, bool PAIR = false, bool BRES = false>
static int launch_tc2(const TcMaps& maps, const TcP& p, int m_tiles, int n_tiles, cudaStream_t s) {
  constexpr size_t stg = 4 * TC2_EPI_GROUPS * 1024 * ((32 * (BN / TC2_EPI_GROUPS) * 2 + 1023) / 1024);
  constexpr size_t smem = STAGES * (TC_BM * TC_BK * 2) + (BRES ? TC2_BRES_MAX_KB : STAGES) * ((PAIR ? BN / 2 : BN) * TC_BK * 2) + stg +
                          (2 * STAGES + 4 + 4 * TC2_EPI_GROUPS + 1) * 8 + 16 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr[S2M2_MAX_DEVICES] = {false};
  static int sms_of[S2M2_MAX_DEVICES] = {0};
  const int di = current_device_index();
  int& sms = sms_of[di];
  if (!attr[di]) {
    cudaError_t e = set_func_attr(conv_tc_persistent_kernel<BN, STAGES, PAIR, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    S2M2_CHECK(e == cudaSuccess, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    attr[di] = true;
  }
  if constexpr (PAIR) {   // clusters of two CTAs (one TPC); one cluster per pair of m-tiles
    const long long pairs = (long long)((m_tiles + 1) / 2) * n_tiles;
    const int clusters = (int)(pairs < sms / 2 ? pairs : sms / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(TC2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled(s) ? 2 : 1;
    if (!plan_mode()) (void)cudaLaunchKernelEx(&cfg, conv_tc_persistent_kernel<BN, STAGES, PAIR, BRES>, maps, p, m_tiles, n_tiles);
    S2M2_LAUNCH_OK("conv_tc_pair");
    return 0;
  }
  const long long tiles = (long long)m_tiles * n_tiles;
  const int grid = (int)(tiles < sms ? tiles : sms);
  launch_k(conv_tc_persistent_kernel<BN, STAGES, PAIR, BRES>, grid, TC2_THREADS, smem, s, maps, p, m_tiles, n_tiles);
  S2M2_LAUNCH_OK("conv_tc_persistent");
  return 0;
}

int conv_tc_launch(const S2m2Conv* a, cudaStream_t s) {
  const S2m2Tensor &x = a->x, &x2 = a->x2, &y = a->y, &res = a->res;
  TcP p{};
  p.bias = a->bias; p.res = a->shuffle2x2 ? nullptr : res.ptr; p.y = y.ptr;
  p.res_ld = res.ptr ? res.ld : 0; p.y_ld = y.ld;
  p.rh = res.ptr ? res.dtype == S2M2_F16 : 0; p.yh = y.dtype == S2M2_F16;
  p.flat = (a->KH == 1 && a->KW == 1) ? 1 : 0;
  p.Cout = a->Cout; p.KW = a->KW; p.ph = a->pad_h; p.pw = a->pad_w; p.taps = a->KH * a->KW; p.stride = a->stride;
  p.cpb1 = x.C / TC_BK; p.cpb2 = x2.ptr ? x2.C / TC_BK : 0;
  p.act = a->act; p.shuffle = a->shuffle2x2; p.Cr = a->shuffle2x2 ? a->Cout / 4 : a->Cout;
  p.emode = a->emode; p.res_from = a->emode == 1 ? a->res_from : 0;
  p.res2 = a->res2.ptr; p.res2_ld = a->res2.ptr ? a->res2.ld : 0;
  long long mtiles;
  if (p.flat && !a->shuffle2x2) {
    p.Ho = 1; p.Wo = (int)((long long)x.N * x.H * x.W); p.tw = TC_BM; p.th = 1; p.tiles_x = cdiv(p.Wo, TC_BM); p.tiles_y = 1;
    mtiles = p.tiles_x;
  } else {
    p.flat = 0;
    p.Ho = x.H / a->stride; p.Wo = x.W / a->stride;
    // patch shape: th*tw = 128, minimise padded-tile waste
    int best_tw = 16;
    long long best = -1;
    for (int tw = 8; tw <= 128; tw *= 2) {
      int th = TC_BM / tw;
      long long tiles = (long long)cdiv(p.Wo, tw) * cdiv(p.Ho, th);
      if (best < 0 || tiles < best || (tiles == best && tw > best_tw)) { best = tiles; best_tw = tw; }
    }
    p.tw = best_tw; p.th = TC_BM / best_tw;
    p.tiles_x = cdiv(p.Wo, p.tw); p.tiles_y = cdiv(p.Ho, p.th);
    mtiles = (long long)p.tiles_x * p.tiles_y * x.N;
  }
  TcMaps maps;
  if (a->stride == 2) {
    for (int q = 0; q < 4; ++q)
      if (int r = encode_act(&maps.act[q], x, false, p.tw, p.th, q)) return r;
  } else {
    if (int r = encode_act(&maps.act[0], x, p.flat, p.tw, p.th)) return r;
    if (x2.ptr) {
      if (int r = encode_act(&maps.act[1], x2, p.flat, p.tw, p.th)) return r;
    } else {
      maps.act[1] = maps.act[0];
    }
    maps.act[2] = maps.act[0];
    maps.act[3] = maps.act[0];
  }
  // N tile: as wide as Cout allows (fewest re-reads of the activations) unless that leaves SMs idle — small
  // problems (1/16, 1/32 resolution) take narrower tiles so that ~every SM gets one.
  int BN = a->Cout >= 256 ? 256 : (a->Cout >= 128 ? 128 : 64);
  while (BN > 64 && mtiles * cdiv(a->Cout, BN) < 120) BN /= 2;
  static const int use_v1 = getenv("S2M2_TC_V1") ? atoi(getenv("S2M2_TC_V1")) : 0;
  // CTA pairs (cta_group::2) for the wide tiles of the deep-K layers (3x3, 3x1/1x3, 5x5: +4..10 % measured) once there is work
  // for every TPC; the K <= 512 linears are epilogue-bound and lose to the pair's extra synchronisation.  S2M2_TC_PAIR=0: off,
  // =2: every BN = 256 layer.
  static const int pair_env = getenv("S2M2_TC_PAIR") ? atoi(getenv("S2M2_TC_PAIR")) : 1;
  const long long nkb_host = (long long)p.taps * (p.cpb1 + p.cpb2);
  static const int pair_min_kb = getenv("S2M2_TC_PAIR_MIN_KB") ? atoi(getenv("S2M2_TC_PAIR_MIN_KB")) : 16;   // 64-channel k-blocks
  const bool pair = pair_env && !use_v1 && BN == 256 && mtiles * cdiv(a->Cout, BN) >= 296 && (nkb_host >= pair_min_kb || pair_env == 2);
  {
    const long long K = (long long)p.taps * (x.C + (x2.ptr ? x2.C : 0));
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)a->Cout}, strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {TC_BK, (cuuint32_t)(pair ? BN / 2 : BN)}, estr[2] = {1, 1};   // a pair's CTA loads half of the rows
    CUresult r = get_encode()(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a->w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  }
  static const int dbg = getenv("S2M2_TC_DBG") ? atoi(getenv("S2M2_TC_DBG")) : 0;
  p.dbg = dbg;
  static const int tile_order = getenv("S2M2_TC_TILE_ORDER") ? atoi(getenv("S2M2_TC_TILE_ORDER")) : 1;   // 1 = m-major, 0 = n-major
  p.m_major = tile_order;
  if (!use_v1) {
    const int nt = cdiv(a->Cout, BN);
    // output tensor map for the TMA-store epilogue: box = {BN/4 channels, the 32 pixels one epilogue warp owns}
    static const int no_tma_store = getenv("S2M2_TC_NO_TMA_STORE") ? atoi(getenv("S2M2_TC_NO_TMA_STORE")) : 0;
    p.tma_store = (!no_tma_store && y.dtype == S2M2_F16 && !a->shuffle2x2) ? 1 : 0;
    if (p.tma_store) {
      const int CG = BN / TC2_EPI_GROUPS;
      cuuint64_t dims[4], strides[3];
      cuuint32_t box[4] = {(cuuint32_t)CG, 32, 1, 1}, estr[4] = {1, 1, 1, 1};
      if (p.flat) {
        dims[0] = y.C; dims[1] = (cuuint64_t)y.N * y.H * y.W; dims[2] = 1; dims[3] = 1;
        strides[0] = (cuuint64_t)y.ld * 2; strides[1] = strides[0] * dims[1]; strides[2] = strides[1];
      } else {
        dims[0] = y.C; dims[1] = y.W; dims[2] = y.H; dims[3] = y.N;
        strides[0] = (cuuint64_t)y.ld * 2; strides[1] = strides[0] * y.W; strides[2] = strides[1] * y.H;
        if (p.tw < 32) { box[1] = p.tw; box[2] = 32 / p.tw; }
      }
      const CUtensorMapSwizzle sw = CG * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (CG * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      CUresult r = get_encode()(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, y.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      S2M2_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(output) failed: %d", (int)r);
      p.tma_res = (p.res && res.dtype == S2M2_F16 && !getenv("S2M2_TC_NO_TMA_RES")) ? 1 : 0;
      if (p.tma_res) {
        cuuint64_t rstr[3];
        if (p.flat) { rstr[0] = (cuuint64_t)res.ld * 2; rstr[1] = rstr[0] * dims[1]; rstr[2] = rstr[1]; }
        else { rstr[0] = (cuuint64_t)res.ld * 2; rstr[1] = rstr[0] * y.W; rstr[2] = rstr[1] * y.H; }
        cuuint64_t rdims[4] = {(cuuint64_t)res.C, dims[1], dims[2], dims[3]};   // (mode 1: the residual has Cout - res_from channels)
        CUresult r2 = get_encode()(&maps.res, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(res.ptr), rdims, rstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        S2M2_CHECK(r2 == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(residual) failed: %d", (int)r2);
      } else {
        maps.res = maps.out;
      }
    } else {
      maps.out = maps.w;
      maps.res = maps.w;
    }
    if (pair) return launch_tc2<256, 5, true>(maps, p, (int)mtiles, nt, s);
    if (BN == 256) return launch_tc2<256, 3>(maps, p, (int)mtiles, nt, s);
    if (BN == 128) return launch_tc2<128, 5>(maps, p, (int)mtiles, nt, s);
    static const int bres_env = getenv("S2M2_TC_BRES") ? atoi(getenv("S2M2_TC_BRES")) : 1;   // 0: stream the weights through the ring
    if (bres_env && nt == 1 && nkb_host <= TC2_BRES_MAX_KB) return launch_tc2<64, 8, false, true>(maps, p, (int)mtiles, nt, s);
    return launch_tc2<64, 8>(maps, p, (int)mtiles, nt, s);
  }
  dim3 grid((unsigned)mtiles, cdiv(a->Cout, BN));
  if (BN == 256) return launch_tc<256, 4>(maps, p, grid, s);
  if (BN == 128) return launch_tc<128, 4>(maps, p, grid, s);
  return launch_tc<64, 4>(maps, p, grid, s);
}

}  // namespace s2m2
