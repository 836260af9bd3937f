// Tensor-core flash attention for fp16 qkv (legacy mma.sync m16n8k16 path, fp32 accumulate / softmax).
// One CTA = 64 queries of one (sequence, head), 4 warps x 16 rows; K/V tiles of 64 keys through shared memory;
// online softmax in registers; P re-used as the A fragment of the PV product.
// Used for every attention without the relative-PE term (attentions.py:50,86,91); the PE variant and head dims
// that are not multiples of 16 / larger than 256 take the CUDA-core kernel in attention.cu.
#include "attention.cuh"

namespace s2m2 {

constexpr int FA_BQ = 64, FA_THREADS = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 -> 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *(uint32_t*)&h;
}

// PE = true: 2-D attention with the contextual relative-PE term (attentions.py:42-48).  Key tiles are then whole
// image rows of the (pe_h x pe_w) token grid (pe_w <= 64), so that
//     pesum[q, 0:16]  = .5 * sum_xj  Mx[q, xj] * pe_x[xq - xj + w - 1],   Mx[q, xj] = sum_yj P[q, (yj, xj)]
//     pesum[q, 16:32] = .5 * sum_yj  My[q, yj] * pe_y[yq - yj + h - 1],   My[q, yj] = sum_xj P[q, (yj, xj)]
// only needs the two marginals of P: Mx accumulates in registers (a thread always owns the same key columns),
// My is one row-sum per tile.  The (N, N, 32) tensor of utils.py:32-60 never exists.
// VS > 1 splits the value / output channels of a head over VS CTAs (each recomputes S = Q K^T): keeps the fp32 output
// accumulator in registers for d = 384 (XL level 0).
// K/V tiles are prefetched NST - 1 tiles ahead (cp.async ring): with ~2 CTAs per SM one tile of lookahead does not cover
// the L2 round trip, two do (1/32-resolution 2-D attention: 44 -> see profiles).
template <int D, int FA_BK, int VS>
constexpr int fa_stages() { return ((64 + 3 * FA_BK) * (D + 8) + 3 * FA_BK * (D / VS + 8)) * 2 <= 120 * 1024 ? 3 : 2; }

template <int D, int FA_BK, bool PE, int VS>
__global__ void __launch_bounds__(FA_THREADS) attention_mma_kernel(AttnP p) {
  pdl_sync();
  constexpr int NST = fa_stages<D, FA_BK, VS>();
  constexpr int LDS = D + 8;               // padded row (halves): conflict-free ldmatrix
  constexpr int KS = D / 16;               // k-steps of QK^T
  constexpr int NG = FA_BK / 8;            // 8-key groups per tile
  constexpr int DV = D / VS;               // value / output channels handled by this CTA
  constexpr int LDV = DV + 8;
  constexpr int OG = DV / 8;               // 8-channel groups of the output
  extern __shared__ __align__(16) __half sm[];
  __half* Qs = sm;                         // [FA_BQ][LDS]
  __half* KV = Qs + FA_BQ * LDS;           // NST stages x { K [FA_BK][LDS], V [FA_BK][LDV] }  (cp.async ring)
  constexpr int STAGE = FA_BK * (LDS + LDV);   // halves per stage
  float* My_raw = (float*)(KV + NST * STAGE);     // PE: [FA_BQ][pe_h] row sums at the running max of that time
  float* M_at = My_raw + FA_BQ * (PE ? p.pe_h : 0);   // PE: [FA_BQ][pe_h] that running max
  float* Mxs = M_at + FA_BQ * (PE ? p.pe_h : 0);      // PE: [FA_BQ][FA_BK] x-marginal (written at the end)
  const int q0 = blockIdx.x * FA_BQ, head = blockIdx.y / VS, vs = blockIdx.y % VS, s = blockIdx.z;
  const int s_kv = (s + p.kv_shift) % p.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __half* base = (const __half*)p.qkv;
  constexpr int CH = D / 8;                // 16-byte chunks per row

  for (int e = tid; e < FA_BQ * CH; e += FA_THREADS) {
    int r = e / CH, c = (e % CH) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < p.N) v = *(const uint4*)(base + ((long long)s * p.N + q0 + r) * p.qkv_ld + head * D + c);
    *(uint4*)(Qs + r * LDS + c) = v;
  }

  float o[OG][4];
#pragma unroll
  for (int i = 0; i < OG; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;   // rows g and g+8 of this warp's 16-row slab
  const float sl2 = p.scale * 1.4426950408889634f;       // softmax in base 2
  const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(Qs + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (lane >> 4) * 8);
  const uint32_t k_addr0 = (uint32_t)__cvta_generic_to_shared(KV + ((lane & 7) + (lane >> 4) * 8) * LDS + ((lane >> 3) & 1) * 8);
  const uint32_t v_addr0 = (uint32_t)__cvta_generic_to_shared(KV + FA_BK * LDS + ((lane & 7) + ((lane >> 3) & 1) * 8) * LDV + (lane >> 4) * 8);
  const uint32_t kv_base = (uint32_t)__cvta_generic_to_shared(KV);

  float mxr[PE ? NG : 1][4];               // PE: x-marginal of P for rows g, g+8 at key columns 8*ng + 2t, +1
#pragma unroll
  for (int i = 0; i < (PE ? NG : 1); ++i) mxr[i][0] = mxr[i][1] = mxr[i][2] = mxr[i][3] = 0.f;
  const int kstep = PE ? p.pe_w : FA_BK;    // keys per tile
  int tile_i = 0;

  // asynchronous K/V tile fetch into stage `st` (rows beyond the tile are zero-filled)
  auto fetch = [&](int k0, int st) {
    const int kend = PE ? k0 + kstep : p.N;
    for (int e = tid; e < FA_BK * CH; e += FA_THREADS) {
      const int r = e / CH, c = (e % CH) * 8;
      const bool ok = k0 + r < kend;
      const __half* src = base + ((long long)s_kv * p.N + (ok ? k0 + r : 0)) * p.qkv_ld + head * D + c;
      cp_async16(kv_base + (uint32_t)(st * STAGE + r * LDS + c) * 2, src + p.C, ok);
      if (c < DV)   // this CTA's slice of the value channels
        cp_async16(kv_base + (uint32_t)(st * STAGE + FA_BK * LDS + r * LDV + c) * 2, src + 2 * p.C + vs * DV, ok);
    }
  };
  // one commit group per tile slot, empty past the last tile, so that "all but the newest NST - 1 groups" is always tile i
  fetch(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  if constexpr (NST == 3) {
    if (kstep < p.N) fetch(kstep, 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  for (int k0 = 0; k0 < p.N; k0 += kstep, ++tile_i) {
    const int kend = PE ? k0 + kstep : p.N;  // first key NOT in this tile
    const int st = tile_i % NST;
    if (k0 + (NST - 1) * kstep < p.N) fetch(k0 + (NST - 1) * kstep, (tile_i + NST - 1) % NST);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if constexpr (NST == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const uint32_t k_addr = k_addr0 + st * STAGE * 2, v_addr = v_addr0 + st * STAGE * 2;

    // ---- S = Q K^T (16 x 64 per warp)
    float sc[NG][4];
#pragma unroll
    for (int i = 0; i < NG; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(q_addr + ks * 32, a0, a1, a2, a3);
#pragma unroll
      for (int ng = 0; ng < NG; ng += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_addr + (ng * 8 * LDS + ks * 16) * 2, b0, b1, b2, b3);
        mma16816(sc[ng], a0, a1, a2, a3, b0, b1);
        mma16816(sc[ng + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // ---- online softmax (rows g = lane/4 and g+8; columns 2t, 2t+1 of each 8-key group)
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int ng = 0; ng < NG; ++ng) {
      int key = k0 + ng * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        bool ok = key + j < kend;
        sc[ng][j] = ok ? sc[ng][j] * sl2 : -1e30f;
        sc[ng][2 + j] = ok ? sc[ng][2 + j] * sl2 : -1e30f;
        mx0 = fmaxf(mx0, sc[ng][j]);
        mx1 = fmaxf(mx1, sc[ng][2 + j]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int ng = 0; ng < NG; ++ng) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sc[ng][j] = exp2f(sc[ng][j] - mn0);
        sc[ng][2 + j] = exp2f(sc[ng][2 + j] - mn1);
        rs0 += sc[ng][j];
        rs1 += sc[ng][2 + j];
      }
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
    if constexpr (PE) {
#pragma unroll
      for (int ng = 0; ng < NG; ++ng) {
        mxr[ng][0] = mxr[ng][0] * c0 + sc[ng][0]; mxr[ng][1] = mxr[ng][1] * c0 + sc[ng][1];
        mxr[ng][2] = mxr[ng][2] * c1 + sc[ng][2]; mxr[ng][3] = mxr[ng][3] * c1 + sc[ng][3];
      }
      float t0 = rs0, t1 = rs1;   // this tile's row sums: y-marginal for image row tile_i
      t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
      t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
      if ((lane & 3) == 0) {
        const int rr0 = warp * 16 + (lane >> 2);
        My_raw[rr0 * p.pe_h + tile_i] = t0;       M_at[rr0 * p.pe_h + tile_i] = mn0;
        My_raw[(rr0 + 8) * p.pe_h + tile_i] = t1; M_at[(rr0 + 8) * p.pe_h + tile_i] = mn1;
      }
    }
#pragma unroll
    for (int i = 0; i < OG; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < FA_BK / 16; ++kk) {
      const uint32_t a0 = pack_h2(sc[2 * kk][0], sc[2 * kk][1]), a1 = pack_h2(sc[2 * kk][2], sc[2 * kk][3]);
      const uint32_t a2 = pack_h2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]), a3 = pack_h2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int og = 0; og < OG; og += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_addr + (kk * 16 * LDV + og * 8) * 2, b0, b1, b2, b3);
        mma16816(o[og], a0, a1, a2, a3, b0, b1);
        mma16816(o[og + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();   // every warp is done with this stage before it is refilled
  }
  // ---- finalise: quad-reduce the row sums, normalise, store fp16
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  if constexpr (PE) {
    // publish the normalised marginals, then every thread produces 16 of the 64 x 32 PE sums of this query tile
    const int lr0 = warp * 16 + (lane >> 2), lr1 = lr0 + 8;
#pragma unroll
    for (int ng = 0; ng < NG; ++ng) {
      const int c = ng * 8 + (lane & 3) * 2;
      Mxs[lr0 * FA_BK + c] = mxr[ng][0] * i0; Mxs[lr0 * FA_BK + c + 1] = mxr[ng][1] * i0;
      Mxs[lr1 * FA_BK + c] = mxr[ng][2] * i1; Mxs[lr1 * FA_BK + c + 1] = mxr[ng][3] * i1;
    }
    if ((lane & 3) == 0) {
      for (int y = 0; y < p.pe_h; ++y) {
        My_raw[lr0 * p.pe_h + y] *= exp2f(M_at[lr0 * p.pe_h + y] - m0) * i0;
        My_raw[lr1 * p.pe_h + y] *= exp2f(M_at[lr1 * p.pe_h + y] - m1) * i1;
      }
    }
    __syncthreads();
    const int ql = tid & 63, part = tid >> 6;      // part 0: x half (channels 0..15), part 1: y half (16..31)
    const int t = q0 + ql;
    if (t < p.N) {
      const int xq = t % p.pe_w, yq = t / p.pe_w;
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
      const int n = part == 0 ? p.pe_w : p.pe_h;
      const float* tab = part == 0 ? p.pe_x : p.pe_y;
      const float* mg = part == 0 ? Mxs + ql * FA_BK : My_raw + ql * p.pe_h;
      const int qc = part == 0 ? xq : yq;
      for (int j = 0; j < n; ++j) {
        const float w = 0.5f * mg[j];
        const float4* row = (const float4*)(tab + (qc - j + n - 1) * 16);
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const float4 tv = __ldg(row + e4);
          acc[4 * e4] = fmaf(w, tv.x, acc[4 * e4]); acc[4 * e4 + 1] = fmaf(w, tv.y, acc[4 * e4 + 1]);
          acc[4 * e4 + 2] = fmaf(w, tv.z, acc[4 * e4 + 2]); acc[4 * e4 + 3] = fmaf(w, tv.w, acc[4 * e4 + 3]);
        }
      }
      __half* ps = (__half*)p.pesum + (((long long)s * p.N + t) * p.heads + head) * 32 + part * 16;
#pragma unroll
      for (int e = 0; e < 16; e += 2) *(__half2*)(ps + e) = __floats2half2_rn(acc[e], acc[e + 1]);
    }
  }
  __half* outp = (__half*)p.out;
#pragma unroll
  for (int og = 0; og < OG; ++og) {
    const int c = head * D + vs * DV + og * 8 + (lane & 3) * 2;
    if (r0 < p.N) *(__half2*)(outp + ((long long)s * p.N + r0) * p.out_ld + c) = __floats2half2_rn(o[og][0] * i0, o[og][1] * i0);
    if (r1 < p.N) *(__half2*)(outp + ((long long)s * p.N + r1) * p.out_ld + c) = __floats2half2_rn(o[og][2] * i1, o[og][3] * i1);
  }
}

template <int D, bool PE, int VS = 1, int BK_OVERRIDE = 0>
static int launch_fa(const AttnP& p, cudaStream_t s) {
  constexpr int FA_BK = BK_OVERRIDE ? BK_OVERRIDE : ((D > 128) ? 32 : 64);   // 32 keeps two CTAs per SM at d = 256
  constexpr int NST = fa_stages<D, FA_BK, VS>();
  static_assert(FA_BQ == 64, "fa_stages() assumes 64-query tiles");
  const size_t smem = (size_t)((FA_BQ + NST * FA_BK) * (D + 8) + NST * FA_BK * (D / VS + 8)) * sizeof(__half) +
                      (PE ? sizeof(float) * FA_BQ * (2 * p.pe_h + FA_BK) : 0);
  S2M2_CHECK(smem <= 200 * 1024, "attention_mma: shared memory %zu too large", smem);
  static bool attr[S2M2_MAX_DEVICES] = {false};
  const int di = current_device_index();
  if (!attr[di]) {
    set_func_attr(attention_mma_kernel<D, FA_BK, PE, VS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr[di] = true;
  }
  dim3 grid(cdiv(p.N, FA_BQ), p.heads * VS, p.S);
  launch_k(attention_mma_kernel<D, FA_BK, PE, VS>, grid, FA_THREADS, smem, s, p);
  S2M2_LAUNCH_OK("attention_mma");
  return 0;
}

int attention_mma_supported(const AttnP& p) {
  if (!p.qh || !p.oh) return 0;
  if (p.pesum) {  // PE variant: fp16 pesum, image rows as key tiles, head dims used at 1/32 resolution
    if (!p.ph_half || p.pe_w > 128 || p.pe_h * p.pe_w != p.N || p.pe_h > 128) return 0;
    if (!(p.d == 16 || p.d == 32 || p.d == 64 || p.d == 96)) return 0;
  }
  if (p.qkv_ld % 8 || p.out_ld % 2 || !aligned(p.qkv, 16) || !aligned(p.out, 4)) return 0;
  switch (p.d) {
    case 16: case 32: case 64: case 96: case 128: case 192: case 256: case 384: return 1;
    default: return 0;
  }
}

int attention_mma_launch(const AttnP& p, cudaStream_t s) {
  if (p.pesum && p.pe_w > 64) {   // wide token grids (XL@2432x2048: 76 columns): one image row = one 128-key tile
    switch (p.d) {
      case 16: return launch_fa<16, true, 1, 128>(p, s);
      case 32: return launch_fa<32, true, 1, 128>(p, s);
      case 64: return launch_fa<64, true, 1, 128>(p, s);
      case 96: return launch_fa<96, true, 1, 128>(p, s);
    }
  }
  if (p.pesum) {
    switch (p.d) {
      case 16: return launch_fa<16, true>(p, s);
      case 32: return launch_fa<32, true>(p, s);
      case 64: return launch_fa<64, true>(p, s);
      case 96: return launch_fa<96, true>(p, s);
    }
  }
  switch (p.d) {
    case 16: return launch_fa<16, false>(p, s);
    case 32: return launch_fa<32, false>(p, s);
    case 64: return launch_fa<64, false>(p, s);
    case 96: return launch_fa<96, false>(p, s);
    case 128: return launch_fa<128, false>(p, s);
    case 192: return launch_fa<192, false>(p, s);
    case 256: return launch_fa<256, false>(p, s);
    case 384: return launch_fa<384, false, 2>(p, s);
  }
  set_error("attention_mma: unsupported head dim %d", p.d);
  return -1;
}

}  // namespace s2m2
