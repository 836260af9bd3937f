// Shared device/host helpers for the s2m2b200 kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/s2m2b200.h"

namespace s2m2 {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern long long g_launches;

#define S2M2_CHECK(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      s2m2::set_error(__VA_ARGS__);           \
      return -1;                              \
    }                                         \
  } while (0)

// ---- plan mode ----------------------------------------------------------------------------------
// With the (thread-local) plan flag set every exported op runs its host-side checks and its dispatch — which kernel takes this
// shape, which return code the caller gets — but enqueues nothing.  The host-side forward (host_forward.cu) walks its schedule
// once in this mode to size the workspace (s2m2b200_workspace_bytes).  On a box without a CUDA device the plan assumes the
// sm_100a target (tests/test_host_forward_cpu.py walks the schedule that way).
extern thread_local int g_plan;
static inline bool plan_mode() { return g_plan != 0; }
bool plan_without_device();
void* plan_encode_stub();   // stands in for cuTensorMapEncodeTiled in a device-less plan: the maps are never used, only the return codes

// ---- op trace -----------------------------------------------------------------------------------
// While a trace is open (s2m2b200_trace_begin / _end, thread-local) every exported op appends one line: its name and its arguments
// (tensor views as [N,H,W,C;ld;dtype;ptr], pointers, scalars).  tests/ compare the op sequence of the C++ host forward with the
// Python schedule's this way, on the CPU, with plan mode on.
bool trace_on();
void trace_append(const char* text);
void trace_fmt_tensor(char* buf, size_t cap, const S2m2Tensor* t);
void trace_fmt_conv(char* buf, size_t cap, const S2m2Conv* c);
struct TraceLine {
  char buf[1024];
  size_t n = 0;
  void put(const char* s) {
    while (*s && n + 1 < sizeof(buf)) buf[n++] = *s++;
    buf[n] = 0;
  }
  void arg(const S2m2Tensor* t) { char b[160]; trace_fmt_tensor(b, sizeof(b), t); put(b); }
  void arg(const S2m2Conv* c) { char b[900]; trace_fmt_conv(b, sizeof(b), c); put(b); }
  void arg(const void* p) { char b[40]; snprintf(b, sizeof(b), " %p", p); put(b); }
  void arg(long long v) { char b[40]; snprintf(b, sizeof(b), " %lld", v); put(b); }
  void arg(int v) { arg((long long)v); }
  void arg(long v) { arg((long long)v); }
  void arg(float v) { char b[40]; snprintf(b, sizeof(b), " %.9g", (double)v); put(b); }
};
template <typename... A>
static inline void trace_op(const char* name, A... a) {
  if (!trace_on()) return;
  TraceLine l;
  l.put(name);
  int dummy[] = {0, (l.arg(a), 0)...};
  (void)dummy;
  trace_append(l.buf);
}
#define S2M2_TRACE(...) s2m2::trace_op(__VA_ARGS__)

#define S2M2_LAUNCH_OK(name)                                                        \
  do {                                                                              \
    if (s2m2::plan_mode()) break;                                                   \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      s2m2::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return -2;                                                                    \
    }                                                                               \
    ++s2m2::g_launches;                                                             \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------------
// Kernels launched with the programmatic-stream-serialization attribute (S2M2_PDL: 0 never, 1 always, default auto = eager
// launches yes, launches under stream capture no; api.cu) may start while the last wave of the previous kernel drains: each kernel
// runs pdl_sync() before it touches global memory, so its CTAs run their shared-memory / TMEM / barrier prologue early, then block
// until that kernel has completed and its writes are visible.  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled(cudaStream_t stream);

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  if (plan_mode()) return;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(stream) ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through cudaGetLastError() in S2M2_LAUNCH_OK
}

// Per-device launch state: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to the (function, device)
// pair, not to the process — one process may run the model on several GPUs (model.to('cuda:1'), PipelinedRunner(device=...)).
constexpr int S2M2_MAX_DEVICES = 64;
static inline int current_device_index() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= S2M2_MAX_DEVICES) d = 0;
  return d;
}

// cudaFuncSetAttribute that a device-less plan (see plan mode above) survives
template <typename F>
static inline cudaError_t set_func_attr(F* kernel, cudaFuncAttribute attr, int value) {
  if (plan_without_device()) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, attr, value);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- dtype-agnostic element access ---------------------------------------------------------------
__device__ __forceinline__ float ld1(const void* p, long long i, int half) {
  return half ? __half2float(((const __half*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st1(void* p, long long i, int half, float v) {
  if (half) ((__half*)p)[i] = __float2half_rn(v);
  else ((float*)p)[i] = v;
}
// 4 consecutive elements; i must be a multiple of 4 and the base suitably aligned (16 B fp32 / 8 B fp16)
__device__ __forceinline__ float4 ld4(const void* p, long long i, int half) {
  if (half) {
    uint2 r = *(const uint2*)((const __half*)p + i);
    __half2 a = *(__half2*)&r.x, b = *(__half2*)&r.y;
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  return *(const float4*)((const float*)p + i);
}
__device__ __forceinline__ void st4(void* p, long long i, int half, float4 v) {
  if (half) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *(uint32_t*)&a;
    r.y = *(uint32_t*)&b;
    *(uint2*)((__half*)p + i) = r;
  } else {
    *(float4*)((float*)p + i) = v;
  }
}

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case S2M2_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case S2M2_ACT_RELU: return fmaxf(x, 0.0f);
    case S2M2_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case S2M2_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// GELU for the fp16 epilogues (the GELU epilogue, not the MMA, bounds the K <= 512 layers):
//   gelu(x) = x * Phi(x),  Phi(x) = 1 / (1 + exp(-2 u(x))),  u = x * P(x^2)   (erf(x / sqrt 2) = tanh(u(x)) defines u; P is a
// degree-4 minimax fit in x^2 — the first two terms are the familiar sqrt(2/pi) (1 + 0.0447 x^2)).  10 instructions (2 MUFU) per
// element instead of ~19 for an erf polynomial; |gelu_fast - gelu| <= 3.5e-6 absolute over all x (fp32 evaluation;
// tests/test_gelu_fit.py), 1e-5 relative for x > 0.01 — 20x below the fp16 rounding of the stored result.  The coefficients
// below are P's multiplied by -2 log2(e) so that exp(-2u) is one ex2.approx.
__device__ __forceinline__ float gelu_fast(float x) {
  const float t = x * x;
  float p = fmaf(-3.2289886e-06f, t, 8.8238130e-05f);
  p = fmaf(p, t, 3.6027454e-04f);
  p = fmaf(p, t, -1.0522669e-01f);
  p = fmaf(p, t, -2.3020454e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));
  return __fdividef(x, 1.0f + e);
}

// activation for fp16 outputs: erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below fp16 rounding) and exp-based
// sigmoid / tanh; the fp32 parity mode keeps act_apply() above
__device__ __forceinline__ float act_apply_fast(float x, int act) {
  switch (act) {
    case S2M2_ACT_GELU: {
      const float z = x * 0.70710678118654752440f, az = fabsf(z);
      const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
      float poly = fmaf(1.061405429f, t, -1.453152027f);
      poly = fmaf(poly, t, 1.421413741f);
      poly = fmaf(poly, t, -0.284496736f);
      poly = fmaf(poly, t, 0.254829592f);
      const float e = copysignf(1.0f - poly * t * __expf(-az * az), z);
      return 0.5f * x * (1.0f + e);
    }
    case S2M2_ACT_RELU: return fmaxf(x, 0.0f);
    case S2M2_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-x));
    case S2M2_ACT_TANH: return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline bool aligned(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }
static inline size_t esize(int dtype) { return dtype == S2M2_F16 ? 2 : 4; }
// can a tensor view be accessed with 4-element vectors?
static inline bool vec4_ok(const S2m2Tensor& t) {
  return t.C % 4 == 0 && t.ld % 4 == 0 && aligned(t.ptr, 4 * esize(t.dtype));
}
static inline long long npix(const S2m2Tensor& t) { return (long long)t.N * t.H * t.W; }

}  // namespace s2m2
