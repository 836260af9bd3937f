// Dust-bin log-Sinkhorn + arg-max + 5-tap soft-argmax with the scan-line's matrix RESIDENT on chip
// (submodules.py:147-243: DispInit._optimal_transport / _sinkhorn / logsumexp_stable and the soft-argmax that follows).
//
// The (w+1)^2 matrix of one scan-line is swept 2*iters + 1 times (three Sinkhorn iterations = six log-sum-exp sweeps, then the
// probabilities).  The first kernel of this stage (match.cu:sinkhorn_kernel) streams it from L2 / HBM on every sweep: at
// L@1216x1024 that is 7 x 95 MB per pair, at XL 7 x 757 MB — the "HBM-bound cost volume" of BASELINE config 4.  Here a
// thread-block CLUSTER owns the scan-line: every CTA keeps a share of the rows of the raw w x w block in its shared memory, the
// cost volume is read from global memory ONCE, every sweep runs out of shared memory, and the only exchange per half-iteration is
// a few hundred floats through distributed shared memory (the CTAs' partial column log-sum-exps, and the row potentials u).
//
//   A(i,j) = cv[i,j], or -1e4 where j > i when use_positivity — exp(-1e4 + ..) underflows to exactly 0 in fp32, in the reference
//   as here, so those entries are neither stored nor visited: row i keeps its columns 0..i only, and rows i and w-1-i share one
//   buffer line of w+1 floats ("pairs").  Pairs are dealt round-robin to the CTAs of the cluster, which balances the triangular
//   work exactly.  Without positivity every CTA keeps full rows i = c (mod CS).
//   Dust-bin row / column = 0 (never stored).  v = log_nu - LSE_i(A + u),  u = log_mu - LSE_j(A + v), v first from u = 0
//   (submodules.py:169-181); online (max, sum) pairs, combined exactly like the max-shifted LSE of submodules.py:147-152.
//
// Warp q' of a CTA owns the CTA's local rows q = q' (mod NW); lanes run over columns j = lane + 32 k  (k < NC registers), so
// shared memory is read with unit stride in both kinds of sweep.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace s2m2 {

// Threads per CTA are a template parameter (512 by default; S2M2_SK_NT=256 selects 8-warp CTAs, up to four per SM).  The sweeps
// are short and every half-iteration ends in block- and cluster-wide barriers after a global-memory load phase: the ncu stall
// samples are barrier / wait, not the exponentials (profiles/r2_dispinit.md).
constexpr float SKC_L2E = 1.4426950408889634f;

__device__ __forceinline__ float skc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (m, s) <- (m, s) (+) t   with s = sum exp(x - m):   one exponential per element
__device__ __forceinline__ void lse_push(float& m, float& s, float t) {
  const float d = t - m;
  const float e = skc_ex2(-fabsf(d) * SKC_L2E);
  s = d > 0.f ? fmaf(s, e, 1.0f) : s + e;
  m = fmaxf(m, t);
}
// (m, s) <- (m, s) (+) (m2, s2)
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float M = fmaxf(m, m2);
  s = s * skc_ex2((m - M) * SKC_L2E) + s2 * skc_ex2((m2 - M) * SKC_L2E);
  m = M;
}

// One row's contribution to the column sums / its own row sum, specialised on the number NCH of 32-column chunks the row really
// has: the rows of the triangular (positivity-masked) matrix are i + 1 entries long, and a loop unrolled to the full width with
// predicates issues (predicated-off) instructions for all of it.
template <int NCH>
__device__ __forceinline__ void skc_col_row(const float* __restrict__ rowp, float ui, const float* mj, float* cs, int lane, int jend) {
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int j = lane + 32 * k;
    if (k < NCH - 1 || j < jend) cs[k] += skc_ex2(((rowp[j] + ui) - mj[k]) * SKC_L2E);
  }
}
template <int NCH>
__device__ __forceinline__ float skc_row_sum(const float* __restrict__ rowp, const float* __restrict__ v, float mi, int lane, int jend) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int j = lane + 32 * k;
    if (k < NCH - 1 || j < jend) {
      const float e = skc_ex2(((rowp[j] + v[j]) - mi) * SKC_L2E);
      if (k & 1) s1 += e; else s0 += e;
    }
  }
  return s0 + s1;
}
#define SKC_DISPATCH(nch, CALL)                                                                                         \
  switch (nch) {                                                                                                          \
    case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break; case 5: CALL(5); break;       \
    case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break; case 9: CALL(9); break; case 10: CALL(10); break;     \
    case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break; case 14: CALL(14); break; case 15: CALL(15); break; \
    case 16: CALL(16); break; case 17: CALL(17); break; case 18: CALL(18); break; case 19: CALL(19); break; default: CALL(20); break; \
  }

struct SkcP {
  const float* cv;
  float *disp, *conf, *occ;
  int* amax;
  int w, pos, iters;
  int lines_ld;      // floats per buffer line: pos ? pad4(w + 1) : pad4(w)
  int nlines_max;    // buffer lines per CTA (ceil over the cluster)
};

// NC = column registers per lane: 32 * NC >= w + 1
template <int NC, int SKC_THREADS>
__global__ void __launch_bounds__(SKC_THREADS, (NC <= 10 ? (SKC_THREADS <= 256 ? 4 : 2) : 1)) sinkhorn_cluster_kernel(const SkcP p) {
  constexpr int SKC_NW = SKC_THREADS / 32;
  constexpr int SKC_SLOTS = SKC_NW / 2;                 // cross-warp combine buffers (two rounds)
  pdl_sync();
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const long long line = blockIdx.x / CS;
  const int w = p.w, W1 = w + 1, pos = p.pos, ld = p.lines_ld;
  extern __shared__ __align__(16) float skc_sh[];
  float* Am = skc_sh;                                  // [nlines_max][ld]
  float* u = Am + (size_t)p.nlines_max * ld;           // [W1] row potentials (all rows; all-gathered through DSMEM)
  float* v = u + W1;                                   // [W1] column potentials
  float* pm = v + W1;                                  // [CS][W1] partial column (max, sum) of every CTA of the cluster
  float* ps = pm + (size_t)CS * W1;
  float* wm = ps + (size_t)CS * W1;                    // [SLOTS][32 * NC] cross-warp combine
  float* ws = wm + SKC_SLOTS * 32 * NC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float log_m_in = logf(1.0f / (2.0f * (float)w));            // log(1/(2w))   (submodules.py:187-189)
  const float log_m_dust = logf((float)w / (2.0f * (float)w));       // log(w/(2w))
  const float log_const = logf(2.0f * (float)w);                      // submodules.py:198-199
  const float* A = p.cv + line * (long long)w * w;

  // local rows.  pos: buffer line l holds pair P = crank + l*CS = rows (P, w-1-P): local row 2l = row P at [0, P],
  // local row 2l+1 = row w-1-P at [P+1, w] (absent when w-1-P == P).  !pos: buffer line l = row crank + l*CS.
  const int npairs = (w + 1) / 2;
  const int nl = pos ? (npairs > crank ? (npairs - crank + CS - 1) / CS : 0) : (w > crank ? (w - crank + CS - 1) / CS : 0);
  const int nrows = pos ? 2 * nl : nl;
  auto row_of = [&](int q, int& i, const float*& rowp, int& jend) -> bool {
    if (pos) {
      const int l = q >> 1, P = crank + l * CS;
      if (q & 1) {
        i = w - 1 - P;
        rowp = Am + (size_t)l * ld + (P + 1);
        jend = i + 1;
        return i != P;
      }
      i = P; rowp = Am + (size_t)l * ld; jend = P + 1;
      return true;
    }
    i = crank + q * CS; rowp = Am + (size_t)q * ld; jend = w;
    return true;
  };

  // ---- the cost volume is read once (only the entries that are ever used)
  for (int q = wid; q < nrows; q += SKC_NW) {
    int i, jend;
    const float* rowp;
    if (!row_of(q, i, rowp, jend)) continue;
    float* dst = const_cast<float*>(rowp);
    const float* src = A + (long long)i * w;
    for (int j = lane; j < jend; j += 32) dst[j] = __ldg(src + j);
  }
  for (int i = threadIdx.x; i < W1; i += SKC_THREADS) { u[i] = 0.f; v[i] = 0.f; }
  __syncthreads();

  // column sweep: v[j] = log_nu[j] - LSE_i (A(i,j) + u[i])  over ALL rows i of the scan-line (+ the dust-bin row, A = 0)
  auto col_sweep = [&]() {
    float cm[NC], cs[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { cm[k] = -3.0e38f; cs[k] = 0.f; }
    for (int q = wid; q < nrows; q += SKC_NW) {
      int i, jend;
      const float* rowp;
      if (!row_of(q, i, rowp, jend)) continue;
      const float ui = u[i];
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (32 * k < jend) {                             // warp-uniform: the chunk intersects this row's columns
          const int j = lane + 32 * k;
          if (j < jend) lse_push(cm[k], cs[k], rowp[j] + ui);
        }
      }
    }
    // the dust-bin column j = w (A = 0 on every row): v[w] = log_nu[w] - LSE_{i = 0..w} u[i].  Every CTA holds all of u and
    // evaluates it for itself, BEFORE the cluster barrier below (after it a faster peer may already be rewriting u).
    if (wid == SKC_NW - 1) {
      float m = -3.0e38f, s = 0.f;
      for (int i = lane; i < W1; i += 32) lse_push(m, s, u[i]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
      }
      if (lane == 0) v[w] = log_m_dust - (m + logf(fmaxf(s, 1e-30f)));
    }
    // combine the warps of this CTA: upper half -> slots, lower half merges and writes back
    if (wid >= SKC_SLOTS) {
#pragma unroll
      for (int k = 0; k < NC; ++k) { wm[(wid - SKC_SLOTS) * 32 * NC + 32 * k + lane] = cm[k]; ws[(wid - SKC_SLOTS) * 32 * NC + 32 * k + lane] = cs[k]; }
    }
    __syncthreads();
    if (wid < SKC_SLOTS) {
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        lse_merge(cm[k], cs[k], wm[wid * 32 * NC + 32 * k + lane], ws[wid * 32 * NC + 32 * k + lane]);
        wm[wid * 32 * NC + 32 * k + lane] = cm[k]; ws[wid * 32 * NC + 32 * k + lane] = cs[k];
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += SKC_THREADS) {
      float m = wm[j], s = ws[j];
#pragma unroll
      for (int ww = 1; ww < SKC_SLOTS; ++ww) lse_merge(m, s, wm[ww * 32 * NC + j], ws[ww * 32 * NC + j]);
      // publish this CTA's partial to every CTA of the cluster (distributed shared memory)
      for (int c = 0; c < CS; ++c) {
        cluster.map_shared_rank(pm, c)[(size_t)crank * W1 + j] = m;
        cluster.map_shared_rank(ps, c)[(size_t)crank * W1 + j] = s;
      }
    }
    cluster.sync();
    for (int j = threadIdx.x; j < w; j += SKC_THREADS) {
      float m = u[w], s = 1.0f;                          // the dust-bin row i = w: A = 0, term = u[w] (written by this CTA only)
      for (int c = 0; c < CS; ++c) lse_merge(m, s, pm[(size_t)c * W1 + j], ps[(size_t)c * W1 + j]);
      v[j] = log_m_in - (m + logf(fmaxf(s, 1e-30f)));
    }
    __syncthreads();
  };

  // ---- bound-shifted sweeps.  After a row update  A(i,j) + u[i] + v[j] <= log_mu[i] <= log(1/2)  for every entry (each term of a
  // normalised row is at most the row's log-mass), so  A(i,j) + u[i] <= log(1/2) - v[j] =: m_j  is a safe shift for column j of the
  // NEXT column sweep — and symmetrically  A(i,j) + v[j] <= log(1/2) - u[i]  for the next row sweep.  With the shift known in
  // advance the log-sum-exp is a single pass of INDEPENDENT exponentials (no running max, no dependent rescaling chain):
  //     LSE = m + log(sum exp(x - m)),   sum in [~1/(2w e), 1]  — far from under- / overflow.
  // Only the very first column sweep (u = v = 0, no bound yet) runs the online (max, sum) version above.
  auto col_sweep_bound = [&]() {
    float cs[NC], mj[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int j = lane + 32 * k;
      cs[k] = 0.f;
      mj[k] = log_m_dust - (j < W1 ? v[j] : 0.f);
    }
    for (int q = wid; q < nrows; q += SKC_NW) {
      int i, jend;
      const float* rowp;
      if (!row_of(q, i, rowp, jend)) continue;
      const float ui = u[i];
      const int nch = (jend + 31) >> 5;                   // warp-uniform
#define SKC_CALL(N) if (N <= NC) skc_col_row<(N <= NC ? N : 1)>(rowp, ui, mj, cs, lane, jend)
      SKC_DISPATCH(nch, SKC_CALL)
#undef SKC_CALL
    }
    // dust-bin column j = w: sum_{i = 0..w} exp(u[i] - m_w), by the last warp of every CTA for itself, before the cluster barrier
    if (wid == SKC_NW - 1) {
      const float mw = log_m_dust - v[w];
      float s0 = 0.f;
      for (int i = lane; i < W1; i += 32) s0 += skc_ex2((u[i] - mw) * SKC_L2E);
      s0 = warp_sum(s0);
      if (lane == 0) v[w] = log_m_dust - (mw + logf(fmaxf(s0, 1e-30f)));
    }
    if (wid >= SKC_SLOTS) {
#pragma unroll
      for (int k = 0; k < NC; ++k) ws[(wid - SKC_SLOTS) * 32 * NC + 32 * k + lane] = cs[k];
    }
    __syncthreads();
    if (wid < SKC_SLOTS) {
#pragma unroll
      for (int k = 0; k < NC; ++k) ws[wid * 32 * NC + 32 * k + lane] += cs[k];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += SKC_THREADS) {
      float sj = ws[j];
#pragma unroll
      for (int ww = 1; ww < SKC_SLOTS; ++ww) sj += ws[ww * 32 * NC + j];
      for (int c = 0; c < CS; ++c) cluster.map_shared_rank(ps, c)[(size_t)crank * W1 + j] = sj;
    }
    cluster.sync();
    for (int j = threadIdx.x; j < w; j += SKC_THREADS) {
      const float m = log_m_dust - v[j];
      float sj = skc_ex2((u[w] - m) * SKC_L2E);          // the dust-bin row i = w: A = 0
      for (int c = 0; c < CS; ++c) sj += ps[(size_t)c * W1 + j];
      v[j] = log_m_in - (m + logf(fmaxf(sj, 1e-30f)));
    }
    __syncthreads();
  };
  auto row_sweep_bound = [&]() {
    for (int q = wid; q < nrows; q += SKC_NW) {
      int i, jend;
      const float* rowp;
      if (!row_of(q, i, rowp, jend)) continue;
      const float mi = log_m_dust - u[i];
      const int nch = (jend + 31) >> 5;
      float sr = 0.f;
#define SKC_CALL(N) if (N <= NC) sr = skc_row_sum<(N <= NC ? N : 1)>(rowp, v, mi, lane, jend)
      SKC_DISPATCH(nch, SKC_CALL)
#undef SKC_CALL
      if (lane == 0) sr += skc_ex2((v[w] - mi) * SKC_L2E);   // dust-bin column: A = 0
      sr = warp_sum(sr);
      if (lane == 0) {
        const float ui = log_m_in - (mi + logf(fmaxf(sr, 1e-30f)));
        for (int c = 0; c < CS; ++c) cluster.map_shared_rank(u, c)[i] = ui;
      }
    }
    if (wid == SKC_NW - 1) {   // dust-bin row i = w: A = 0 everywhere
      const float mw = log_m_dust - u[w];
      float s0 = 0.f;
      for (int j = lane; j < W1; j += 32) s0 += skc_ex2((v[j] - mw) * SKC_L2E);
      s0 = warp_sum(s0);
      if (lane == 0) u[w] = log_m_dust - (mw + logf(fmaxf(s0, 1e-30f)));
    }
    cluster.sync();
  };

  // (the first column sweep runs with u = 0: v = log_nu - LSE_i A, submodules.py:172)
  col_sweep();
  row_sweep_bound();        // u = 0 so far: A(i,j) + v[j] <= log_nu[j] - 0 holds after the first column normalisation
  for (int it = 1; it < p.iters; ++it) {
    col_sweep_bound();
    row_sweep_bound();
  }

  // ---- probabilities of this CTA's rows: row sum (occ), first arg-max, 5 taps around it (submodules.py:198-241)
  for (int q = wid; q < nrows; q += SKC_NW) {
    int i, jend;                                        // jend: masked entries are exactly 0 (submodules.py:223)
    const float* row;
    if (!row_of(q, i, row, jend)) continue;
    const float ui = u[i] + log_const;
    float best = -1.f, sum = 0.f;
    int bidx = 0x7fffffff;
    for (int j = lane; j < jend; j += 32) {
      const float pj = expf(row[j] + ui + v[j]);
      sum += pj;
      if (pj > best) { best = pj; bidx = j; }
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    float tap = 0.f;
    {
      const int j = bidx - 2 + lane;
      if (lane < 5 && j >= 0 && j < jend) tap = expf(row[j] + ui + v[j]);
    }
    float cf = 0.f, cr = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float wk = __shfl_sync(0xffffffffu, tap, k);
      cf += wk;
      cr += wk * (float)(bidx + k - 2);
    }
    if (lane == 0) {
      cr = (cr + 1e-4f) / (cf + 1e-4f);
      const long long o = line * w + i;
      p.disp[o] = (float)i - cr;
      p.conf[o] = cf;
      p.occ[o] = sum;
      p.amax[o] = bidx;
    }
  }
  cluster.sync();   // no CTA may exit while a peer could still write into its shared memory
}

static void skc_geometry(int w, int pos, int CS, int& ld, int& nlines) {
  if (pos) {
    ld = (w + 1 + 3) & ~3;
    nlines = ((w + 1) / 2 + CS - 1) / CS;
  } else {
    ld = (w + 3) & ~3;
    nlines = (w + CS - 1) / CS;
  }
}
static size_t skc_smem(int w, int pos, int CS, int NC, int NT) {
  int ld, nlines;
  skc_geometry(w, pos, CS, ld, nlines);
  return sizeof(float) * ((size_t)nlines * ld + 2 * (size_t)(w + 1) + 2 * (size_t)CS * (w + 1) + 2 * (size_t)(NT / 64) * 32 * NC);
}

// 0 = launched; 1 = this width is not taken by the cluster kernel (the caller runs the streaming kernel); < 0 = error
int sinkhorn_cluster_launch(const float* cv, int B, int h, int w, float* disp, float* conf, float* occ, int* amax, int pos, int iters,
                            cudaStream_t s) {
  if (w + 1 > 640 || w < 8) return 1;
  const int NC = (w + 1 + 31) / 32 <= 10 ? 10 : 20;
  static const int cs_env = getenv("S2M2_SK_CS") ? atoi(getenv("S2M2_SK_CS")) : 0;
  static const int nt_env = getenv("S2M2_SK_NT") ? atoi(getenv("S2M2_SK_NT")) : 0;
  const int NT = (NC == 10 && nt_env == 256) ? 256 : 512;      // measured at L, batch 8: 512 threads x 4-CTA clusters 1.29 ms, 256 threads 1.38-1.41 ms
  constexpr size_t SMEM_MAX = 220 * 1024;
  // the smallest cluster whose share lets at least `want` CTAs live on one SM (227 KB): more co-resident CTAs hide the barriers
  const size_t want_bytes = NT == 256 ? 56 * 1024 : 112 * 1024;
  int CS = 0;
  if ((cs_env == 1 || cs_env == 2 || cs_env == 4 || cs_env == 8) && skc_smem(w, pos, cs_env, NC, NT) <= SMEM_MAX) CS = cs_env;
  if (!CS)
    for (int c = 1; c <= 8 && !CS; c *= 2)
      if (skc_smem(w, pos, c, NC, NT) <= want_bytes) CS = c;
  if (!CS)
    for (int c = 1; c <= 8 && !CS; c *= 2)
      if (skc_smem(w, pos, c, NC, NT) <= SMEM_MAX) CS = c;
  if (!CS) return 1;
  const size_t smem = skc_smem(w, pos, CS, NC, NT);
  SkcP p{};
  p.cv = cv; p.disp = disp; p.conf = conf; p.occ = occ; p.amax = amax; p.w = w; p.pos = pos; p.iters = iters;
  skc_geometry(w, pos, CS, p.lines_ld, p.nlines_max);
  void (*kernel)(const SkcP) = NC == 10 ? (NT == 256 ? sinkhorn_cluster_kernel<10, 256> : sinkhorn_cluster_kernel<10, 512>)
                                        : sinkhorn_cluster_kernel<20, 512>;
  const int ki = NC == 10 ? (NT == 256 ? 0 : 1) : 2;
  static bool attr[S2M2_MAX_DEVICES][3] = {{false}};
  const int di = current_device_index();
  if (!attr[di][ki]) {
    cudaError_t e = set_func_attr(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    S2M2_CHECK(e == cudaSuccess, "sinkhorn_cluster: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr[di][ki] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((long long)B * h * CS));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled(s) ? 2 : 1;
  if (!plan_mode()) (void)cudaLaunchKernelEx(&cfg, kernel, p);
  S2M2_LAUNCH_OK("sinkhorn_cluster");
  return 0;
}

}  // namespace s2m2
