// common.cuh — shared host/device helpers for libdcsnet_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/dcsnet.h"

namespace dcs {

// ---- error plumbing (include/dcsnet.h: 0 = ok, <0 argument error, >0 cudaError_t)
char* err_buf();
int set_error(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define DCS_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::dcs::set_error(-1, __VA_ARGS__);       \
  } while (0)

#define DCS_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::dcs::set_error((int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// every kernel launch goes through this so dcs_launch_count() is an honest count
#define DCS_LAUNCHED()                                                                  \
  do {                                                                                  \
    ::dcs::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
    cudaError_t e__ = cudaPeekAtLastError();                                            \
    if (e__ != cudaSuccess)                                                             \
      return ::dcs::set_error((int)e__, "kernel launch failed: %s", cudaGetErrorString(e__)); \
  } while (0)

static inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- element access for the two activation storage types (complex element = 2 scalars)
template <typename T> struct Elem;
template <> struct Elem<float> {
  using pair_t = float2;
  static __device__ __forceinline__ float2 ldc(const float* p, int64_t i) { return reinterpret_cast<const float2*>(p)[i]; }
  static __device__ __forceinline__ void stc(float* p, int64_t i, float2 v) { reinterpret_cast<float2*>(p)[i] = v; }
};
template <> struct Elem<__nv_bfloat16> {
  using pair_t = __nv_bfloat162;
  static __device__ __forceinline__ float2 ldc(const __nv_bfloat16* p, int64_t i) {
    return __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(p)[i]);
  }
  static __device__ __forceinline__ void stc(__nv_bfloat16* p, int64_t i, float2 v) {
    reinterpret_cast<__nv_bfloat162*>(p)[i] = __float22bfloat162_rn(v);
  }
};

// Packed fp32x2 FMA (sm_100: FFMA2): acc.{x,y} = a.{x,y} * b.{x,y} + acc.{x,y}, two IEEE fma.rn in ONE issue slot.
// The CUDA-core kernels here are issue-bound (FFMA + LDS share the schedulers), so halving the FMA instruction count is
// the lever; results are bit-identical to two scalar fmaf.
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  acc = *reinterpret_cast<float2*>(&rc);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
// Short-latency forms for the serial LSTM recurrence and the mask tail: ex2.approx-based exp (rel. error ~2^-22
// over the argument range that matters, |x| < 88) and a correctly rounded reciprocal; absolute error of the results
// is <= ~2e-7, far inside the 1e-5 parity budget, at ~1/6 of the instruction count of expf/tanhf + IEEE division.
__device__ __forceinline__ float fast_sigmoid(float v) { return __frcp_rn(1.f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float v) {
  const float e = __expf(-2.f * fabsf(v));             // in (0, 1]: no overflow
  const float t = (1.f - e) * __frcp_rn(1.f + e);
  return copysignf(t, v);
}
// Shortest dependent chains for the serial LSTM cell update: MUFU.EX2 + MUFU.RCP (approximate reciprocal, ~1 ulp) with
// no Newton fix-up or special-case branch; absolute error <= ~3e-7.
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));   // bare MUFU.RCP (no range-scaling code as in __fdividef)
  return r;
}
__device__ __forceinline__ float quick_sigmoid(float v) { return rcp_approx(1.f + __expf(-v)); }
__device__ __forceinline__ float quick_tanh(float v) { return 2.f * rcp_approx(1.f + __expf(-2.f * v)) - 1.f; }
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == DCS_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DCS_ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  if (act == DCS_ACT_SIGMOID) return sigmoidf_(v);
  return v;
}

// bound_cRM (network_functions.py:77-88): tanh(|m|) e^{j atan2(im, re+eps)} applied twice inside (both atan2s add eps)
__device__ __forceinline__ float2 bound_crm_dev(float2 m, float eps, bool exact) {
  if (!exact) {
    // tanh(|m|) e^{j atan2(im, re+eps)} twice, in the algebraic form (cos(atan2(y,x)) = x/hypot(x,y)) with
    // short-latency tanh / rsqrt: ~25 instructions instead of ~250 for the literal transcendental sequence
    const float mag2 = m.x * m.x + m.y * m.y;
    const float t = fast_tanh(mag2 * rsqrtf(fmaxf(mag2, 1e-37f)));
    const float x1 = m.x + eps, q1 = x1 * x1 + m.y * m.y;
    float r1, i1;
    if (q1 == 0.f) { r1 = t; i1 = 0.f; } else { const float s = t * rsqrtf(q1); r1 = x1 * s; i1 = m.y * s; }
    const float x2 = r1 + eps, q2 = x2 * x2 + i1 * i1;
    if (q2 == 0.f) return make_float2(t, 0.f);
    const float s2 = t * rsqrtf(q2);
    return make_float2(x2 * s2, i1 * s2);
  }
  const float t = tanhf(sqrtf(m.x * m.x + m.y * m.y));
  if (exact) {
    const float th1 = atan2f(m.y, m.x + eps);
    const float r1 = t * cosf(th1), i1 = t * sinf(th1);
    const float th2 = atan2f(i1, r1 + eps);
    return make_float2(t * cosf(th2), t * sinf(th2));
  }
  float x1 = m.x + eps, h1 = sqrtf(x1 * x1 + m.y * m.y);
  float r1, i1;
  if (h1 == 0.f) { r1 = t; i1 = 0.f; } else { const float s = t / h1; r1 = x1 * s; i1 = m.y * s; }
  float x2 = r1 + eps, h2 = sqrtf(x2 * x2 + i1 * i1);
  if (h2 == 0.f) return make_float2(t, 0.f);
  const float s2 = t / h2;
  return make_float2(x2 * s2, i1 * s2);
}

// bound_cRM for the bf16 / tensor-core tail (cconv_strip.cu): the same algebraic form with bare MUFU instructions only
// (ex2.approx, rcp.approx, rsqrt.approx — no range fix-up, no Newton step, no slow-path call): ~22 instead of ~38
// instructions per call; abs. error ~3e-7 on a value bounded by 1, against the bf16 inputs' 4e-3.
__device__ __forceinline__ float2 bound_crm_mufu(float2 m, float eps) {
  const float mag2 = m.x * m.x + m.y * m.y;
  float rs, e, r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(mag2, 1e-37f)));
  const float mag = mag2 * rs;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.8853900817779268f * mag));      // e^(-2 |m|) in (0, 1]
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  const float t = (1.f - e) * r;                                                        // tanh(|m|) >= 0
  const float x1 = m.x + eps, q1 = x1 * x1 + m.y * m.y;
  float s1;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(q1));
  const bool z1 = q1 == 0.f;
  const float r1 = z1 ? t : x1 * (t * s1), i1 = z1 ? 0.f : m.y * (t * s1);
  const float x2 = r1 + eps, q2 = x2 * x2 + i1 * i1;
  float s2;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s2) : "f"(q2));
  const bool z2 = q2 == 0.f;
  return make_float2(z2 ? t : x2 * (t * s2), z2 ? 0.f : i1 * (t * s2));
}

}  // namespace dcs
