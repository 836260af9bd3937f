// C-ABI glue: error state, conv dispatch (CUDA-core vs tcgen05), host-side PE table.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include <cuda.h>

#include "common.cuh"

namespace s2m2 {

static thread_local char g_err[512] = "";
long long g_launches = 0;
thread_local int g_plan = 0;

bool plan_without_device() {
  if (!g_plan) return false;
  static int have = -1;
  if (have < 0) {
    int n = 0;
    have = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0;
    (void)cudaGetLastError();
  }
  return have == 0;
}

int pdl_mode() {   // S2M2_PDL: 0 = never, 1 = always, unset / 2 = auto: on for launches into a stream that is not being captured
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("S2M2_PDL");
    mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }
  return mode;
}
// Programmatic dependent launch: every kernel runs pdl_sync() before it touches global memory, so with the attribute its CTAs are
// scheduled (and run their shared-memory / TMEM / barrier prologue) while the previous kernel drains.  Eager launches gain from it
// (the ~880 launches of a single-pair forward: 32.0 -> 31.1 ms against 30.6 ms for the graph replay); inside a captured CUDA graph the
// kernel-to-kernel gaps are already hidden and it measured no gain (34.1 vs 34.3 ms in round 1), so "auto" leaves captures alone.
bool pdl_enabled(cudaStream_t stream) {
  const int m = pdl_mode();
  if (m != 2) return m == 1;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return st == cudaStreamCaptureStatusNone;
}

static thread_local bool g_trace_on = false;
static thread_local char* g_trace = nullptr;
static thread_local size_t g_trace_len = 0, g_trace_cap = 0;
bool trace_on() { return g_trace_on; }
void trace_append(const char* text) {
  if (!g_trace_on) return;
  const size_t n = strlen(text);
  if (g_trace_len + n + 2 > g_trace_cap) {
    g_trace_cap = (g_trace_cap ? g_trace_cap * 2 : (1 << 20)) + n + 2;
    g_trace = (char*)realloc(g_trace, g_trace_cap);
  }
  memcpy(g_trace + g_trace_len, text, n);
  g_trace_len += n;
  g_trace[g_trace_len++] = '\n';
  g_trace[g_trace_len] = 0;
}
void trace_fmt_tensor(char* buf, size_t cap, const S2m2Tensor* t) {
  if (!t || !t->ptr) snprintf(buf, cap, " [null]");
  else snprintf(buf, cap, " [%d,%d,%d,%d;ld%lld;%s;%p]", t->N, t->H, t->W, t->C, (long long)t->ld, t->dtype == S2M2_F16 ? "f16" : "f32", t->ptr);
}
void trace_fmt_conv(char* buf, size_t cap, const S2m2Conv* c) {
  char x[160], x2[160], y[160], r[160], r2[160];
  trace_fmt_tensor(x, sizeof(x), &c->x); trace_fmt_tensor(x2, sizeof(x2), &c->x2); trace_fmt_tensor(y, sizeof(y), &c->y);
  trace_fmt_tensor(r, sizeof(r), &c->res); trace_fmt_tensor(r2, sizeof(r2), &c->res2);
  snprintf(buf, cap, " x%s x2%s y%s res%s res2%s w %p %s bias %p k%dx%d s%d p%d,%d cout%d act%d shuf%d batch%d bs%lld,%lld,%lld impl%d emode%d from%d",
           x, x2, y, r, r2, c->w, c->w_dtype == S2M2_F16 ? "f16" : "f32", (const void*)c->bias, c->KH, c->KW, c->stride, c->pad_h, c->pad_w,
           c->Cout, c->act, c->shuffle2x2, c->batch, (long long)c->x_bs, (long long)c->w_bs, (long long)c->y_bs, c->impl, c->emode, c->res_from);
}

static CUresult encode_stub(CUtensorMap* m, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  memset(m, 0, sizeof(*m));
  return CUDA_SUCCESS;
}
void* plan_encode_stub() { return (void*)encode_stub; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int conv_simt_launch(const S2m2Conv* a, cudaStream_t s);
int conv_tc_supported(const S2m2Conv* a);
int conv_tc_launch(const S2m2Conv* a, cudaStream_t s);

}  // namespace s2m2

using namespace s2m2;

extern "C" {

const char* s2m2b200_last_error(void) { return g_err; }
int s2m2b200_trace_begin(void) {
  g_trace_len = 0;
  if (g_trace) g_trace[0] = 0;
  g_trace_on = true;
  return 0;
}
void s2m2b200_trace_note(const char* text) { trace_append(text); }
int64_t s2m2b200_trace_end(char* buf, int64_t cap) {   // stops the trace; copies min(len, cap-1) bytes + NUL; returns the full length
  g_trace_on = false;
  if (buf && cap > 0) {
    const size_t n = g_trace_len < (size_t)cap - 1 ? g_trace_len : (size_t)cap - 1;
    if (n) memcpy(buf, g_trace, n);
    buf[n] = 0;
  }
  return (int64_t)g_trace_len;
}
int s2m2b200_set_plan_mode(int on) {   // returns the previous value (thread-local)
  const int prev = g_plan;
  g_plan = on ? 1 : 0;
  return prev;
}
const char* s2m2b200_version(void) { return "s2m2b200 0.2 sm_100a (tcgen05 f16 implicit GEMM, fused token kernels, split-fp16 cost volume; cuda-core fp32 parity path)"; }
int64_t s2m2b200_launch_count(void) { return g_launches; }

int s2m2b200_conv2d(const S2m2Conv* a, void* stream) {
  S2M2_TRACE("conv2d", a);
  const S2m2Tensor &x = a->x, &y = a->y;
  S2M2_CHECK(x.ptr && y.ptr && a->w, "conv2d: null pointer");
  S2M2_CHECK(a->KH > 0 && a->KW > 0 && a->stride > 0 && a->Cout > 0, "conv2d: bad geometry");
  const int Ho = (x.H + 2 * a->pad_h - a->KH) / a->stride + 1, Wo = (x.W + 2 * a->pad_w - a->KW) / a->stride + 1;
  if (a->x2.ptr) {
    S2M2_CHECK(a->x2.N == x.N && a->x2.H == x.H && a->x2.W == x.W && a->x2.dtype == x.dtype, "conv2d: x2 must match x");
  }
  if (a->shuffle2x2) {
    S2M2_CHECK(a->Cout % 4 == 0 && y.C == a->Cout / 4 && y.H == 2 * Ho && y.W == 2 * Wo && y.N == x.N && !a->res.ptr,
               "conv2d(shuffle): y must be (N,2Ho,2Wo,Cout/4) and no residual");
  } else {
    S2M2_CHECK(y.C == a->Cout && y.H == Ho && y.W == Wo && y.N == x.N, "conv2d: y shape (%d,%d,%d,%d) != expected (%d,%d,%d,%d)",
               y.N, y.H, y.W, y.C, x.N, Ho, Wo, a->Cout);
    if (a->res.ptr) S2M2_CHECK(a->res.C == y.C - (a->emode == 1 ? a->res_from : 0) && npix(a->res) == npix(y), "conv2d: residual shape mismatch");
    if (a->res2.ptr) S2M2_CHECK(a->res2.C == y.C && npix(a->res2) == npix(y) && a->res2.dtype == S2M2_F16, "conv2d: res2 shape / dtype mismatch");
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (a->emode != 0) {          // fused gate / GRU epilogues: tcgen05 path only (the caller composes them from separate ops otherwise)
    S2M2_CHECK(a->emode >= 1 && a->emode <= 3 && a->res.ptr && !a->shuffle2x2, "conv2d: bad epilogue mode %d", a->emode);
    S2M2_CHECK(a->emode == 1 || a->res2.ptr, "conv2d: epilogue mode %d needs res2", a->emode);
    if (a->impl == 1 || !conv_tc_supported(a)) return 1;   // not applicable: the caller composes the step from separate ops
    return conv_tc_launch(a, s);
  }
  if (a->impl == 2) {
    S2M2_CHECK(conv_tc_supported(a), "conv2d: tcgen05 kernel forced but shape/dtype unsupported");
    return conv_tc_launch(a, s);
  }
  if (a->impl == 0 && conv_tc_supported(a)) return conv_tc_launch(a, s);
  return conv_simt_launch(a, s);
}

// utils.py:32-60 — one axis of the relative positional table.  torch.linspace semantics: the first half
// is start + i*step, the second half end - (n-1-i)*step.
static float linspace_at(float start, float end, int n, int i) {
  if (n == 1) return start;
  float step = (end - start) / (float)(n - 1);
  return i < n / 2 ? start + (float)i * step : end - (float)(n - 1 - i) * step;
}

int s2m2b200_pe_table(int32_t n, float* host_out) {
  S2M2_CHECK(n > 0 && host_out, "pe_table: bad arguments");
  const int L = 2 * n + 1, D = 16;
  const float sig = 5.0f / 32.0f;
  for (int r = 0; r < L; ++r) {
    float pos = tanhf(linspace_at(-3.f, 3.f, L, r));
    float row[16];
    float nrm = 0.f;
    for (int c = 0; c < D; ++c) {
      float x = (linspace_at(-1.f, 1.f, D, c) - pos) / sig;
      float v = fabsf(x) < 1e-6f ? 1.0f : sinf(3.1415f * x) / (3.1415f * x);  // literal 3.1415 as in utils.py:6
      row[c] = v;
      nrm += v * v;
    }
    nrm = fmaxf(sqrtf(nrm), 1e-12f);  // F.normalize eps
    for (int c = 0; c < D; ++c) host_out[r * D + c] = row[c] / nrm;
  }
  return 0;
}

}  // extern "C"
