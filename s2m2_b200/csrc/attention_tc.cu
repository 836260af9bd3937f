// tcgen05 whole-row attention for the 1-D (scan-line) blocks: every key of a sequence fits on chip (N <= 384),
// so the softmax is exact in one pass over TMEM — no online rescaling.
//
// Persistent CTAs (one per SM) walk over work tiles = 128 queries of one (sequence, head):
//     QK    S[128 x N]   = Q K^T      tcgen05.mma (M=128, N=64 per key block, K=d), S in TMEM columns [0, N_pad)
//     SM    P = softmax(S)             16 warps: 4 per TMEM lane quarter, each owning every 4th 32-column chunk of its 32
//                                      rows (partial max / sum exchanged through shared memory); fp16 P is written to
//                                      shared memory in the K-major 128B-swizzled layout the tensor core reads as A
//     PV    O[128 x d]   = P V         tcgen05.mma (M=128, N=d, K=64 per key block); V is consumed as it lies in memory
//                                      ([key][channel] = MN-major B operand, b_major = 1): no transpose pass
//     EPI   out = O / rowsum           TMEM -> registers -> global (full-sector stores)
// The tile loop is software-pipelined: Q has its own buffer (P does not alias it) and O lives in TMEM columns
// [o_col, o_col + d) next to S, so the tensor pipe runs   PV(i) -> QK(i+1)   back to back while the softmax warps drain
// O(i); only the key blocks of S(i+1) whose columns overlap O(i) (d = 256, N > 256) wait for that drain.  K/V stream
// through one TMA ring that runs ahead across tile boundaries, so the next tile's loads hide behind the current math.
//   warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..17 = softmax / epilogue.
// Reference op: F.scaled_dot_product_attention in the 1-D blocks (attentions.py:50,86,91,119-161).
#include <stdlib.h>

#include "tc_common.cuh"

#include "attention.cuh"

namespace s2m2 {
using namespace tc;

constexpr int ATC_SPLIT = 4;                        // softmax warps per TMEM lane quarter (columns are dealt round-robin)
constexpr int ATC_SM_THREADS = 128 * ATC_SPLIT;
constexpr int ATC_THREADS = 64 + ATC_SM_THREADS;

struct AtcP {
  void* out;
  long long out_ld;
  int S, N, heads, C, kv_shift, N_pad, NKB, o_col, q_tiles, tiles;
  float sl2;  // softmax scale * log2(e)
};
struct AtcMaps {
  CUtensorMap q;   // qkv (tokens x 3C): box {64 ch, 128 tokens}
  CUtensorMap k;   // qkv (tokens x 3C): box {64 ch, 64 tokens}
};

template <int D>
constexpr int atc_stages() { return D >= 256 ? 2 : 8; }   // ring slots of 64 keys x D channels; as many bytes in flight as fit

__device__ __forceinline__ float ex2_approx(float x) {   // x <= 0 here: no range fix-up needed, denormal results flush to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static int atc_nst_override() {   // d = 256: a three-slot K/V ring where it fits next to Q and P (N <= 192; 430 -> 394 us at
  static const int v = getenv("S2M2_ATC_NST") ? atoi(getenv("S2M2_ATC_NST")) : 3;   // S=4096, N=176); S2M2_ATC_NST=2 keeps two slots
  return v;
}
static size_t atc_smem_bytes(int D, int NKB) {
  int NST = D >= 256 ? 2 : 8;
  if (D >= 256 && atc_nst_override() == 3 && NKB <= 3) NST = 3;
  return (size_t)128 * D * 2 + (size_t)NKB * 16384 + (size_t)NST * 64 * D * 2 + (6 + 2 * NST) * 8 + 16 + 2 * ATC_SPLIT * 128 * 4 + 1024;
}

template <int D, int NSTX = 0>
__global__ void __launch_bounds__(ATC_THREADS, 1) attention_tc_kernel(const __grid_constant__ AtcMaps maps, const AtcP p) {
  constexpr int KC = D / 64;                       // 64-channel chunks of the head dim
  constexpr uint32_t Q_BYTES = 128 * D * 2;        // KC chunks of [128 rows][128 B]
  constexpr uint32_t BLK_BYTES = 64 * D * 2;       // one ring slot: K block or V block, both KC x [64 keys][128 B]
  constexpr int NST = NSTX ? NSTX : atc_stages<D>();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sP = smem + Q_BYTES;                    // NKB key blocks x [128 rows][128 B]
  uint8_t* sRing = sP + (uint32_t)p.NKB * 16384u;
  uint64_t* bars = (uint64_t*)(sRing + NST * BLK_BYTES);   // qfull, qfree, sfull, pready, ofull, ofree, full[NST], empty[NST]
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 + 2 * NST);
  float* sRed = (float*)(tmem_slot + 4);           // [2][ATC_SPLIT][128]: partial row max, partial row sum
  const uint32_t qfull = smem_u32(bars), qfree = qfull + 8, sfull = qfull + 16, pready = qfull + 24, ofull = qfull + 32, ofree = qfull + 40;
  const uint32_t full0 = qfull + 48, empty0 = full0 + 8 * NST;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NKB = p.NKB;

  if (threadIdx.x == 0) {
    mbar_init(qfull, 1);
    mbar_init(qfree, 1);
    mbar_init(sfull, 1);
    mbar_init(pready, ATC_SM_THREADS);
    mbar_init(ofull, 1);
    mbar_init(ofree, ATC_SM_THREADS);
    for (int i = 0; i < NST; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue above overlaps the previous kernel's tail; global memory is touched only below

  // tile -> (query tile, head, sequence): the query tiles of one scan-line are adjacent, so they run on different SMs at
  // the same time and share its K/V through L2
  auto decode = [&](int tile, int& q0, int& head, int& s) {
    q0 = (tile % p.q_tiles) * 128;
    const int r = tile / p.q_tiles;
    head = r % p.heads;
    s = r / p.heads;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t rit = 0, lt = 0;
      auto ring_load = [&](int col, int row) {
        const int st = rit % NST;
        mbar_wait(empty0 + 8 * st, ((rit / NST) & 1) ^ 1);
        mbar_expect_tx(full0 + 8 * st, BLK_BYTES);
        const uint32_t dst = smem_u32(sRing + st * BLK_BYTES);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(&maps.k, dst + kc * 8192, full0 + 8 * st, col + kc * 64, row);
        ++rit;
      };
      auto load_q = [&](int tile) {
        int q0, head, s;
        decode(tile, q0, head, s);
        mbar_expect_tx(qfull, Q_BYTES);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(&maps.q, smem_u32(sQ + kc * 16384), qfull, head * D + kc * 64, s * p.N + q0);
      };
      if ((int)blockIdx.x < p.tiles) load_q(blockIdx.x);
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        int q0, head, s;
        decode(tile, q0, head, s);
        const int s_kv = (s + p.kv_shift) % p.S;
        for (int b = 0; b < NKB; ++b) ring_load(p.C + head * D, s_kv * p.N + b * 64);
        if (tile + (int)gridDim.x < p.tiles) {       // Q of the next tile as soon as QK^T of this one has read the buffer
          mbar_wait(qfree, lt & 1);
          load_q(tile + gridDim.x);
        }
        for (int b = 0; b < NKB; ++b) ring_load(2 * p.C + head * D, s_kv * p.N + b * 64);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // O = P V: B = V block as loaded, [64 keys][64 ch] x KC chunks = MN-major (channels contiguous): b_major bit 16
      const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t rit = 0;
      // S = Q K^T for the tile with local index lt
      auto issue_qk = [&](uint32_t lt) {
        mbar_wait(qfull, lt & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool drained = (lt == 0);
        for (int b = 0; b < NKB; ++b, ++rit) {
          if (!drained && 64 * (b + 1) > p.o_col) {   // these S columns still hold O of the previous tile
            mbar_wait(ofree, (lt - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            drained = true;
          }
          const int st = rit % NST;
          mbar_wait(full0 + 8 * st, (rit / NST) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            const uint64_t ad = make_sdesc(smem_u32(sQ + kc * 16384)), bd = make_sdesc(smem_u32(sRing + st * BLK_BYTES + kc * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + b * 64, ad + 2 * k, bd + 2 * k, idesc_s, (kc | k) != 0);
          }
          umma_commit(empty0 + 8 * st);
        }
        umma_commit(sfull);
        umma_commit(qfree);
      };
      uint32_t lt = 0;
      if ((int)blockIdx.x < p.tiles) issue_qk(0);
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        mbar_wait(pready, lt & 1);                   // P is in shared memory, S has been read out of TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int b = 0; b < NKB; ++b, ++rit) {       // O += P[:, 64 b ..] V_blk
          const int st = rit % NST;
          mbar_wait(full0 + 8 * st, (rit / NST) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ad = make_sdesc(smem_u32(sP + b * 16384)), bd = make_sdesc_mn(smem_u32(sRing + st * BLK_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // A: +32 B per 16 keys within the 128 B row; B: +16 key rows = 2 swizzle atoms = 2048 B
            umma_f16(tmem_base + p.o_col, ad + 2 * k, bd + (2048 >> 4) * k, idesc_o, (b | k) != 0);
          umma_commit(empty0 + 8 * st);
        }
        umma_commit(ofull);
        if (tile + (int)gridDim.x < p.tiles) issue_qk(lt + 1);
      }
    }
  } else {
    // ================= softmax + epilogue: thread = one query row x every ATC_SPLIT-th column chunk =================
    const int q = warp & 3, h = (warp - 2) >> 2;   // TMEM lane quarter (fixed by warp id % 4), column split
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool odd = lane & 1;
    __half* ob = (__half*)p.out;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
      int q0, head, s;
      decode(tile, q0, head, s);
      const int t = q0 + r;
      const bool valid = t < p.N;
      mbar_wait(sfull, lt & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // pass 1: row max over the valid keys (32 columns per TMEM load: fewer load/wait round trips)
      float m = -3.0e38f;
      for (int c0 = 32 * h; c0 < p.N_pad; c0 += 32 * ATC_SPLIT) {
        uint32_t rr[32];
        tmem_ld32(trow + c0, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.N) m = fmaxf(m, __uint_as_float(rr[i]));
      }
      sRed[h * 128 + r] = m;
      asm volatile("bar.sync 1, %0;" ::"n"(ATC_SM_THREADS) : "memory");
#pragma unroll
      for (int j = 0; j < ATC_SPLIT; ++j) m = fmaxf(m, sRed[j * 128 + r]);
      // pass 2: p = exp2((s - m) * scale*log2e), row sum, fp16 P into the swizzled A-operand layout
      float l = 0.f;
      const float msc = m * p.sl2;
      for (int c0 = 32 * h; c0 < p.N_pad; c0 += 32 * ATC_SPLIT) {
        uint32_t rr[32];
        tmem_ld32(trow + c0, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __half2 h[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a = (c0 + i < p.N) ? ex2_approx(fmaf(__uint_as_float(rr[i]), p.sl2, -msc)) : 0.f;
          const float b = (c0 + i + 1 < p.N) ? ex2_approx(fmaf(__uint_as_float(rr[i + 1]), p.sl2, -msc)) : 0.f;
          l += a + b;
          h[i >> 1] = __floats2half2_rn(a, b);
        }
        uint8_t* blk = sP + (c0 >> 6) * 16384 + r * 128;               // key block, this row
        const int u0 = (c0 & 63) >> 3;                                  // first 16-byte unit of this chunk within the 128 B row
#pragma unroll
        for (int u = 0; u < 4; ++u) *(uint4*)(blk + (((u0 + u) ^ (r & 7)) << 4)) = *(uint4*)&h[4 * u];
      }
      sRed[(ATC_SPLIT + h) * 128 + r] = l;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy P writes -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // our TMEM reads are ordered before the next MMAs
      mbar_arrive(pready);
      asm volatile("bar.sync 1, %0;" ::"n"(ATC_SM_THREADS) : "memory");
      l = 0.f;
#pragma unroll
      for (int j = 0; j < ATC_SPLIT; ++j) l += sRed[(ATC_SPLIT + j) * 128 + r];

      // epilogue: O / l -> fp16 -> global, full 32 B sectors per store instruction via lane-pair exchange
      mbar_wait(ofull, lt & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float inv = 1.f / l;
      const bool pvalid = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
      const long long off0 = ((long long)s * p.N + t) * p.out_ld + head * D;
      const long long poff0 = __shfl_xor_sync(0xffffffffu, off0, 1);
#pragma unroll 1
      for (int c0 = 16 * h; c0 < D; c0 += 16 * ATC_SPLIT) {
        uint32_t rr[16];
        tmem_ld16(trow + p.o_col + c0, rr);
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
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // O has been read: its columns may be overwritten
      mbar_arrive(ofree);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host ----------------------------------------------------------------------------------------
int attention_tc_supported(const AttnP& p) {
  if (!p.qh || !p.oh || p.pesum) return 0;
  if (!(p.d == 64 || p.d == 128 || p.d == 256)) return 0;
  if (p.N < 48 || p.N > 384) return 0;
  if (p.qkv_ld % 8 || p.out_ld % 8 || !aligned(p.qkv, 16) || !aligned(p.out, 16)) return 0;
  if (atc_smem_bytes(p.d, (p.N + 63) / 64) > 227 * 1024) return 0;     // d = 256 with N > 320: the mma.sync kernel takes it
  return tc::device_ok() ? 1 : 0;
}

size_t attention_tc_workspace_bytes(const AttnP& p) { (void)p; return 16; }   // no scratch needed any more (kept for the ABI)

template <int D, int NSTX = 0>
static int launch_atc(const AtcMaps& maps, const AtcP& p, cudaStream_t s) {
  const size_t smem = atc_smem_bytes(D, p.NKB);
  S2M2_CHECK(smem <= 227 * 1024, "attention_tc: shared memory %zu too large", smem);
  static bool attr[S2M2_MAX_DEVICES] = {false};
  static int sms_of[S2M2_MAX_DEVICES] = {0};
  const int di = current_device_index();
  int& sms = sms_of[di];
  if (!attr[di]) {
    cudaError_t e = set_func_attr(attention_tc_kernel<D, NSTX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    S2M2_CHECK(e == cudaSuccess, "attention_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    attr[di] = true;
  }
  launch_k(attention_tc_kernel<D, NSTX>, p.tiles < sms ? p.tiles : sms, ATC_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("attention_tc");
  return 0;
}

int attention_tc_launch(const AttnP& a, void* workspace, cudaStream_t s) {
  AtcP p{};
  const int D = a.d;
  p.out = a.out; p.out_ld = a.out_ld; p.S = a.S; p.N = a.N; p.heads = a.heads; p.C = a.C; p.kv_shift = a.kv_shift;
  p.N_pad = (a.N + 63) / 64 * 64; p.NKB = p.N_pad / 64;
  p.o_col = (p.N_pad + D <= 512) ? p.N_pad : 512 - D;      // O beside S when both fit, else overlapping S's last key blocks
  p.q_tiles = (a.N + 127) / 128;
  p.tiles = p.q_tiles * a.heads * a.S;
  p.sl2 = a.scale * 1.4426950408889634f;
  (void)workspace;
  AtcMaps maps;
  const long long tokens = (long long)a.S * a.N;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(3 * a.C), (cuuint64_t)tokens}, strides[1] = {(cuuint64_t)a.qkv_ld * 2};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t boxq[2] = {64, 128}, boxk[2] = {64, 64};
    CUresult r1 = tc::encode_fn()(&maps.q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.qkv), dims, strides, boxq, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = tc::encode_fn()(&maps.k, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a.qkv), dims, strides, boxk, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, "attention_tc: cuTensorMapEncodeTiled(qkv) failed: %d %d", (int)r1, (int)r2);
  }
  if (D == 256 && atc_nst_override() == 3 && p.NKB <= 3) return launch_atc<256, 3>(maps, p, s);
  if (D == 256) return launch_atc<256>(maps, p, s);
  if (D == 128) return launch_atc<128>(maps, p, s);
  return launch_atc<64>(maps, p, s);
}

}  // namespace s2m2
