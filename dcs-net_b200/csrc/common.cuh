// common.cuh — shared host/device helpers for libdcsnet_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

static inline int num_sms() {   // of the CURRENT device (cached per device: one process may drive several GPUs)
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
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
template <> struct Elem<__half> {
  using pair_t = __half2;
  static __device__ __forceinline__ float2 ldc(const __half* p, int64_t i) {
    return __half22float2(reinterpret_cast<const __half2*>(p)[i]);
  }
  static __device__ __forceinline__ void stc(__half* p, int64_t i, float2 v);
};

// ---- 16-bit activation storage (the tensor-core modes): DCS_F16 = IEEE half (11-bit significand, the default: every
// stored activation is rounded once per layer, and bf16's 8 bits put the enhanced spectrogram at 3.6e-3 of its maximum on
// the randomised-BN parity state against the 2e-3 budget; fp16 gives 4e-4) or DCS_BF16 (wide range, for checkpoints whose
// activations exceed fp16's 65504).  fp16 stores SATURATE to +-65504 instead of producing inf.
// pack (lo, hi) -> one 32-bit word of two 16-bit values, lo in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <typename T> __device__ __forceinline__ uint32_t pack_h2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack_h2<__nv_bfloat16>(float lo, float hi) { return pack_bf16x2(lo, hi); }
template <> __device__ __forceinline__ uint32_t pack_h2<__half>(float lo, float hi) { return pack_f16x2(lo, hi); }
// runtime-selected form for kernels that are not templated on the storage type (uniform branch)
__device__ __forceinline__ uint32_t pack_h2_rt(float lo, float hi, bool f16) { return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
// unpack one 32-bit word of two 16-bit values -> (lo, hi) fp32
template <typename T> __device__ __forceinline__ float2 unpack_h2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack_h2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack_h2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_float<__half>(float v) {
  const uint32_t w = pack_f16x2(v, 0.f);
  return __ushort_as_half((unsigned short)(w & 0xffffu));
}
template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
__device__ __forceinline__ void Elem<__half>::stc(__half* p, int64_t i, float2 v) {
  reinterpret_cast<uint32_t*>(p)[i] = pack_f16x2(v.x, v.y);
}
// runtime-typed loads of saved activations (the training step's backward kernels read them in fp32 or 16-bit storage; uniform branch)
__device__ __forceinline__ float2 ld_c(const void* p, int64_t i, int dt) {            // complex element i
  if (dt == DCS_F32) return reinterpret_cast<const float2*>(p)[i];
  const uint32_t w = reinterpret_cast<const uint32_t*>(p)[i];
  return dt == DCS_F16 ? unpack_h2<__half>(w) : unpack_h2<__nv_bfloat16>(w);
}
__device__ __forceinline__ float4 ld_c2(const void* p, int64_t i, int dt) {           // complex elements 2i, 2i + 1
  if (dt == DCS_F32) return reinterpret_cast<const float4*>(p)[i];
  const uint2 w = reinterpret_cast<const uint2*>(p)[i];
  const float2 a = dt == DCS_F16 ? unpack_h2<__half>(w.x) : unpack_h2<__nv_bfloat16>(w.x);
  const float2 b = dt == DCS_F16 ? unpack_h2<__half>(w.y) : unpack_h2<__nv_bfloat16>(w.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
static inline bool is_h16(int dtype) { return dtype == DCS_BF16 || dtype == DCS_F16; }
static inline bool is_dtype(int dtype) { return dtype == DCS_F32 || dtype == DCS_BF16 || dtype == DCS_F16; }

// ---- pooled sums (numerators of ComplexAdaptiveAvgPool2d(1)): accumulated with 64-bit INTEGER atomics in fixed point
// (Q35.28), so the result does not depend on the order in which CTAs / warps arrive: two runs on the same input are
// bit-identical (float atomics made the channel gates, and through them every later layer, differ from run to run by
// ~1e-3 of the bf16 mode's tolerance budget).  Each partial (a warp's / CTA's fp32 sum of >= 32 values, computed in a
// fixed order) is rounded once to 2^-28 = 3.7e-9; range +-3.4e10.
constexpr float kPoolScale = 268435456.f, kPoolInvScale = 1.f / 268435456.f;
__device__ __forceinline__ void pool_add(long long* acc, float partial) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)__float2ll_rn(partial * kPoolScale));
}
__device__ __forceinline__ float pool_mean(const long long* acc, int64_t i, float inv_hw) {
  return __ll2float_rn(acc[i]) * (kPoolInvScale * inv_hw);
}
// The same accumulators in MAX mode (the real path's AdaptiveMaxPool2d(1), r_network.py:11,23): a value is stored as its
// order-preserving unsigned image + 1, so that the memset-zero state means "empty" and atomicMax is exact and
// order-independent.
__device__ __forceinline__ unsigned long long pool_max_encode(float v) {
  const uint32_t b = __float_as_uint(v);
  return (unsigned long long)((b & 0x80000000u) ? ~b : (b | 0x80000000u)) + 1ull;
}
__device__ __forceinline__ void pool_max(long long* acc, float partial) {
  atomicMax(reinterpret_cast<unsigned long long*>(acc), pool_max_encode(partial));
}
__device__ __forceinline__ float pool_max_value(const long long* acc, int64_t i) {
  const unsigned long long e = (unsigned long long)acc[i];
  if (e == 0ull) return -INFINITY;
  const uint32_t o = (uint32_t)(e - 1ull);
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

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
