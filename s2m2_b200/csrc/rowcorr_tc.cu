// All-pairs scan-line correlation (the cost volume of DispInit, submodules.py:216-217) on tcgen05 with fp32-grade accuracy:
//     cv[b, y, i, j] = sum_c LN(f)[b, y, i, c] * LN(f)[B + b, y, j, c]          (LN = the affine layer_norm of submodules.py:165,216)
// The arg-max over this volume must not move (north_star: bit-exact arg-max index), so fp16 operands are not enough and the
// round-1 path ran it as an fp32 CUDA-core GEMM (23 TFLOP/s, 4.3 ms per 8 pairs).  Here every normalised feature x is split into
// two fp16 numbers  x = hi + lo  (hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits together) and the product is evaluated as
//     hi.hi + hi.lo + lo.hi        (three tcgen05.mma chains into ONE fp32 accumulator in TMEM; the dropped lo.lo term is 2^-22)
// — the accuracy of an fp32 dot product (whose own summation noise is larger) at tensor-core rate.
//
//   kernel 1  ln_split_kernel      LayerNorm (fp32 statistics, affine) of the fp16 transformer features -> two fp16 planes (hi, lo)
//   kernel 2  rowcorr_tc_kernel    one CTA = 128 left pixels of one scan-line x all right pixels of that line:
//             A (this tile's hi / lo planes, all C channels) resident in shared memory, the right image's hi / lo planes stream
//             through a TMA ring in [128 pixels][64 channels] boxes; column chunks of 128 right pixels complete one after the
//             other, so the epilogue (TMEM -> swizzled staging -> cp.async.bulk.tensor store of fp32 rows) of chunk c overlaps the
//             MMAs of chunk c+1.  Out-of-range rows / columns are clipped by the store's tensor map.
//   warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..5 = epilogue (one per TMEM lane quarter).
#include <stdlib.h>

#include "tc_common.cuh"

namespace s2m2 {
using namespace tc;

// ---- kernel 1: LayerNorm + hi/lo split -----------------------------------------------------------------------------
// 8 lanes per pixel (4 pixels per warp in flight: the statistics need 3 shuffle steps and every lane keeps up to four 16-byte
// loads outstanding); C % 64 == 0, C <= 256.  planes: [2][npix][C] fp16 (hi plane, then lo plane)
__global__ void __launch_bounds__(256) ln_split_kernel(const __half* __restrict__ x, long long x_ld, long long npix, int C,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       __half* __restrict__ planes) {
  pdl_sync();
  const int sub = threadIdx.x & 7;                           // lane within the pixel's group of 8
  const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long ng = ((long long)gridDim.x * blockDim.x) >> 3;
  const int nt = C >> 6;                                      // 16-byte units per lane: C / 8 units over 8 lanes
  const long long last = npix - 1;
  for (long long pg = g0; pg < ((npix + 3) / 4) * 4; pg += ng) {   // whole warps stay in the loop (shuffles below)
    const bool live = pg <= last;
    const long long pix = live ? pg : last;
    const __half* row = x + pix * x_ld;
    float v[4][8];
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < nt) {
        const uint4 raw = *(const uint4*)(row + (sub + 8 * t) * 8);
        const __half2* h = (const __half2*)&raw;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[t][2 * i] = f.x; v[t][2 * i + 1] = f.y; s += f.x + f.y; }
      }
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < nt) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[t][i] - mean; q += d * d; }
      }
    q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
    const float rstd = rsqrtf(q / (float)C + 1e-5f);
    if (live) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < nt) {
          const int un = sub + 8 * t;
          const float4 g0v = __ldg((const float4*)(gamma + un * 8)), g1v = __ldg((const float4*)(gamma + un * 8 + 4));
          const float4 b0v = __ldg((const float4*)(beta + un * 8)), b1v = __ldg((const float4*)(beta + un * 8 + 4));
          const float gg[8] = {g0v.x, g0v.y, g0v.z, g0v.w, g1v.x, g1v.y, g1v.z, g1v.w};
          const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
          __half2 hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = (v[t][2 * i] - mean) * rstd * gg[2 * i] + bb[2 * i];
            const float b = (v[t][2 * i + 1] - mean) * rstd * gg[2 * i + 1] + bb[2 * i + 1];
            const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
            hi[i] = __halves2half2(ah, bh);
            lo[i] = __floats2half2_rn(a - __half2float(ah), b - __half2float(bh));
          }
          *(uint4*)(planes + pix * C + un * 8) = *(uint4*)&hi[0];
          *(uint4*)(planes + (npix + pix) * C + un * 8) = *(uint4*)&lo[0];
        }
    }
  }
}

// ---- kernel 2 ----------------------------------------------------------------------------------------------------------
constexpr int RC_THREADS = 64 + 128;
constexpr int RC_NB = 5;                          // ring entries of [128 right pixels][64 channels] = 16 KB
constexpr uint32_t RC_ENT = 16384;

struct RcP {
  int w, h, B, KC, NCH;                           // KC = C / 64 k-blocks, NCH = ceil(w / 128) column chunks
  long long npix;                                 // pixels per plane (2B * h * w)
};
struct RcMaps {
  CUtensorMap planes;                             // fp16 [2 * npix][C]: box {64 ch, 128 pixels}
  CUtensorMap cv;                                 // fp32 {w cols, w rows, B*h lines}: box {32, 32, 1}, 128B swizzle
};

__device__ __forceinline__ void rc_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int KC>
__global__ void __launch_bounds__(RC_THREADS, 1) rowcorr_tc_kernel(const __grid_constant__ RcMaps maps, const RcP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                      // [2 planes][KC] x [128 rows][128 B]
  uint8_t* sB = sA + 2 * KC * RC_ENT;                      // ring
  uint8_t* sStage = sB + RC_NB * RC_ENT;                   // 4 epilogue warps x [32 rows][128 B]
  uint64_t* bars = (uint64_t*)(sStage + 4 * 4096);         // afull, cfull[4], full[NB], empty[NB]
  uint32_t* tmem_slot = (uint32_t*)(bars + 5 + 2 * RC_NB);
  const uint32_t afull = smem_u32(bars), cfull0 = afull + 8, full0 = afull + 40, empty0 = full0 + 8 * RC_NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, lineidx = blockIdx.y;         // M tile of the scan-line, scan-line = b * h + y
  const int b = lineidx / p.h, y = lineidx - b * p.h;
  const long long rowA = ((long long)b * p.h + y) * p.w + mt * 128;                 // left image pixels
  const long long rowB = ((long long)(p.B + b) * p.h + y) * p.w;                    // right image scan-line
  const int NCH = p.NCH;

  if (threadIdx.x == 0) {
    mbar_init(afull, 1);
    for (int c = 0; c < 4; ++c) mbar_init(cfull0 + 8 * c, 1);
    for (int i = 0; i < RC_NB; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
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
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(afull, 2 * KC * RC_ENT);
      for (int pl = 0; pl < 2; ++pl)
        for (int kb = 0; kb < KC; ++kb)
          tma_load_2d(&maps.planes, smem_u32(sA + (pl * KC + kb) * RC_ENT), afull, kb * 64, (int)(pl * p.npix + rowA));
      uint32_t rit = 0;
      for (int c = 0; c < NCH; ++c)
        for (int kb = 0; kb < KC; ++kb)
          for (int pl = 0; pl < 2; ++pl, ++rit) {
            const int st = rit % RC_NB;
            mbar_wait(empty0 + 8 * st, ((rit / RC_NB) & 1) ^ 1);
            mbar_expect_tx(full0 + 8 * st, RC_ENT);
            tma_load_2d(&maps.planes, smem_u32(sB + st * RC_ENT), full0 + 8 * st, kb * 64, (int)(pl * p.npix + rowB + c * 128));
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // M = 128, N = 128, fp16 -> fp32
      mbar_wait(afull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t rit = 0;
      for (int c = 0; c < NCH; ++c) {
        const uint32_t tacc = tmem_base + c * 128;
        for (int kb = 0; kb < KC; ++kb) {
          const uint64_t ahi = make_sdesc(smem_u32(sA + kb * RC_ENT)), alo = make_sdesc(smem_u32(sA + (KC + kb) * RC_ENT));
          {   // right hi plane: hi.hi + lo.hi
            const int st = rit % RC_NB;
            mbar_wait(full0 + 8 * st, (rit / RC_NB) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t bd = make_sdesc(smem_u32(sB + st * RC_ENT));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tacc, ahi + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tacc, alo + 2 * k, bd + 2 * k, idesc, 1);
            umma_commit(empty0 + 8 * st);
            ++rit;
          }
          {   // right lo plane: hi.lo
            const int st = rit % RC_NB;
            mbar_wait(full0 + 8 * st, (rit / RC_NB) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t bd = make_sdesc(smem_u32(sB + st * RC_ENT));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tacc, ahi + 2 * k, bd + 2 * k, idesc, 1);
            umma_commit(empty0 + 8 * st);
            ++rit;
          }
        }
        umma_commit(cfull0 + 8 * c);
      }
    }
  } else {
    // ================= epilogue: warp % 4 = TMEM lane quarter = 32 rows of the tile =================
    const int q = warp & 3;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* stg = sStage + q * 4096;
    const uint32_t stg_u32 = smem_u32(stg);
    const int row0 = mt * 128 + q * 32;                    // first left pixel of this warp's rows
    for (int c = 0; c < NCH; ++c) {
      mbar_wait(cfull0 + 8 * c, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int cc = 0; cc < 4; ++cc) {                     // 32-column pieces of the chunk
        const int col0 = c * 128 + cc * 32;
        if (col0 >= p.w || row0 >= p.w) break;             // warp-uniform
        uint32_t rr[32];
        tmem_ld32(trow + c * 128 + cc * 32, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the staging tile
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 8; ++u)                        // row = lane: 128 B, 16-byte units XOR-swizzled by (row & 7)
          *(uint4*)(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) = make_uint4(rr[4 * u], rr[4 * u + 1], rr[4 * u + 2], rr[4 * u + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          rc_tma_store_3d(&maps.cv, stg_u32, col0, row0, lineidx);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int KC>
static int launch_rc(const RcMaps& maps, const RcP& p, cudaStream_t s) {
  constexpr size_t smem = (size_t)2 * KC * RC_ENT + RC_NB * RC_ENT + 4 * 4096 + (5 + 2 * RC_NB) * 8 + 16 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr[S2M2_MAX_DEVICES] = {false};
  const int di = current_device_index();
  if (!attr[di]) {
    cudaError_t e = set_func_attr(rowcorr_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    S2M2_CHECK(e == cudaSuccess, "rowcorr_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr[di] = true;
  }
  launch_k(rowcorr_tc_kernel<KC>, dim3((p.w + 127) / 128, p.B * p.h), RC_THREADS, smem, s, maps, p);
  S2M2_LAUNCH_OK("rowcorr_tc");
  return 0;
}

}  // namespace s2m2

using namespace s2m2;

// Scratch bytes for the hi / lo planes of s2m2b200_row_corr (0 = shape not taken by the tcgen05 path).
extern "C" int64_t s2m2b200_row_corr_workspace_bytes(int32_t B, int32_t h, int32_t w, int32_t C) {
  if (C % 64 || C > 256 || w % 4 || w < 32 || w > 512) return 0;   // (ln_split: C % 64; rowcorr: <= 4 k-blocks, <= 4 column chunks)
  return (int64_t)2 * 2 * B * h * w * C * 2;
}

// cv[b,y,i,j] = <LN(f)[b,y,i,:], LN(f)[B+b,y,j,:]> with the affine LayerNorm fused in front (submodules.py:165,216-217).
// f: (2B, h, w, C) fp16 NHWC view (left images first), gamma / beta: fp32 (C), cv: (B, h, w, w) fp32 dense, work: scratch of
// s2m2b200_row_corr_workspace_bytes.  Returns 0 = launched, 1 = not applicable (caller: layernorm + the fp32 CUDA-core GEMM).
extern "C" int s2m2b200_row_corr(const S2m2Tensor* f, const float* gamma, const float* beta, float* cv, void* work, int64_t work_bytes,
                                 void* stream) {
  S2M2_TRACE("row_corr", f, gamma, beta, cv, work, work_bytes);
  const int C = f->C, w = f->W, h = f->H, B = f->N / 2;
  if (f->dtype != S2M2_F16 || f->N % 2 || !work) return 1;
  const int64_t need = s2m2b200_row_corr_workspace_bytes(B, h, w, C);
  if (need == 0 || work_bytes < need) return 1;
  if (f->ld % 8 || !aligned(f->ptr, 16) || !aligned(work, 128) || !aligned(cv, 16) || !gamma || !beta || !aligned(gamma, 16) || !aligned(beta, 16)) return 1;
  if (!tc::device_ok()) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  const long long npix = (long long)f->N * h * w;
  {
    const long long blocks_needed = (npix + 31) / 32;                 // 32 pixels per 256-thread block and iteration
    const int blocks = (int)(blocks_needed < 148 * 16 ? blocks_needed : 148 * 16);
    launch_k(ln_split_kernel, blocks, 256, 0, s, (const __half*)f->ptr, (long long)f->ld, npix, C, gamma, beta, (__half*)work);
    S2M2_LAUNCH_OK("ln_split");
  }
  RcP p{};
  p.w = w; p.h = h; p.B = B; p.KC = C / 64; p.NCH = (w + 127) / 128; p.npix = npix;
  RcMaps maps;
  cuuint32_t estr[3] = {1, 1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(2 * npix)}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 128};
    CUresult r = tc::encode_fn()(&maps.planes, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, work, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "row_corr: cuTensorMapEncodeTiled(planes) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)w, (cuuint64_t)B * h}, strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * w * 4};
    cuuint32_t box[3] = {32, 32, 1};
    CUresult r = tc::encode_fn()(&maps.cv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, cv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    S2M2_CHECK(r == CUDA_SUCCESS, "row_corr: cuTensorMapEncodeTiled(cv) failed: %d", (int)r);
  }
  switch (p.KC) {
    case 1: return launch_rc<1>(maps, p, s);
    case 2: return launch_rc<2>(maps, p, s);
    case 3: return launch_rc<3>(maps, p, s);
    default: return launch_rc<4>(maps, p, s);
  }
}
