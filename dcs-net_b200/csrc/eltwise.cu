// eltwise.cu — bandwidth-bound element-wise stages: folded eval-mode ComplexBatchNorm2d (+activation), the
// bound_cRM x2 / mask (.) Y / subtraction tail, dtype conversion, library bookkeeping.
//
// Reference call sites: ComplexBatchNorm2d (complexPyTorch 0.3) at /root/reference/c_network.py:101,113,148;
// bound_cRM network_functions.py:77-88 (called at c_network.py:225 and again at network_functions.py:394);
// complex_mat_mult network_functions.py:90-96; combine network_functions.py:396-397 (dcs), 434 (dc).
#include <stdarg.h>
#include "common.cuh"

namespace dcs {

std::atomic<uint64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

template <typename TI, typename TO>
__global__ void cbn_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ aff, int64_t n, int C,
                           int act) {
  // n = total complex elements (pixels * C)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float* a = aff + 6 * c;
    const float2 v = Elem<TI>::ldc(x, i);
    const float re = act_apply(a[0] * v.x + a[1] * v.y + a[4], act);
    const float im = act_apply(a[2] * v.x + a[3] * v.y + a[5], act);
    Elem<TO>::stc(y, i, make_float2(re, im));
  }
}

// bound_cRM (network_functions.py:77-88).  `exact` evaluates the reference's transcendental sequence literally.
__device__ __forceinline__ float2 bound_crm(float2 m, float eps, bool exact) {
  const float t = tanhf(sqrtf(m.x * m.x + m.y * m.y));
  if (exact) {
    const float th1 = atan2f(m.y, m.x + eps);
    const float r1 = t * cosf(th1), i1 = t * sinf(th1);
    const float th2 = atan2f(i1, r1 + eps);
    return make_float2(t * cosf(th2), t * sinf(th2));
  }
  // cos(atan2(y, x)) = x / hypot(x, y), sin(atan2(y, x)) = y / hypot(x, y);  atan2(0, 0) = 0
  float x1 = m.x + eps, h1 = sqrtf(x1 * x1 + m.y * m.y);
  float r1, i1;
  if (h1 == 0.f) { r1 = t; i1 = 0.f; } else { const float s = t / h1; r1 = x1 * s; i1 = m.y * s; }
  float x2 = r1 + eps, h2 = sqrtf(x2 * x2 + i1 * i1);
  if (h2 == 0.f) return make_float2(t, 0.f);
  const float s2 = t / h2;
  return make_float2(x2 * s2, i1 * s2);
}

__global__ void mask_combine_kernel(const dcs_mask_combine_params p) {
  const float2* raw = reinterpret_cast<const float2*>(p.net_raw);
  const float2* Y = reinterpret_cast<const float2*>(p.noisy_spec);
  float2* o_net = reinterpret_cast<float2*>(p.net_out);
  float2* o_mask = reinterpret_cast<float2*>(p.mask);
  float2* o_noise = reinterpret_cast<float2*>(p.noise_spec);
  float2* o_clean = reinterpret_cast<float2*>(p.clean_spec);
  const bool exact = p.exact_polar != 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 m1 = bound_crm(__ldg(raw + i), p.atan2_eps, exact);  // C_NETWORK.forward's own bound (c_network.py:225)
    const float2 m2 = bound_crm(m1, p.atan2_eps, exact);              // the step function's bound (network_functions.py:394)
    const float2 y = __ldg(Y + i);
    const float2 prod = cmul(y, m2);
    if (o_net) o_net[i] = m1;
    if (o_mask) o_mask[i] = m2;
    if (p.combine == DCS_COMBINE_DCS) {
      if (o_noise) o_noise[i] = prod;
      if (o_clean) o_clean[i] = make_float2(y.x - prod.x, y.y - prod.y);
    } else {
      if (o_clean) o_clean[i] = prod;
    }
  }
}

__global__ void bound_crm_kernel(const float2* __restrict__ x, float2* __restrict__ y, int64_t n, float eps, int exact) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = bound_crm(__ldg(x + i), eps, exact != 0);
}
__global__ void cmul_kernel(const float2* __restrict__ a, const float2* __restrict__ b, float2* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = cmul(__ldg(a + i), __ldg(b + i));
}
// cRM (network_functions.py:62-75): M = S * conj(Y) / (|Y|^2 + eps)
__global__ void crm_kernel(const float2* __restrict__ s, const float2* __restrict__ yn, float2* __restrict__ m, int64_t n, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 S = __ldg(s + i), Y = __ldg(yn + i);
    const float den = Y.x * Y.x + Y.y * Y.y + eps;
    m[i] = make_float2((Y.x * S.x + Y.y * S.y) / den, (Y.x * S.y - Y.y * S.x) / den);
  }
}
template <typename T>
__global__ void upsample_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int uh, int uw) {
  const int64_t n = (int64_t)B * H * uh * W * uw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ox = (int)(r % (W * uw)); r /= (W * uw);
    const int oy = (int)(r % (H * uh));
    const int b = (int)(r / (H * uh));
    Elem<T>::stc(y, i, Elem<T>::ldc(x, (((int64_t)b * H + oy / uh) * W + ox / uw) * C + c));
  }
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ s, TO* __restrict__ d, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    Elem<TO>::stc(d, i, Elem<TI>::ldc(s, i));
}

static inline int ew_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace dcs

using namespace dcs;

// every (input, output) storage pair of the element-wise kernels
#define DCS_DISPATCH_IO(IN, OUT, F)                                             \
  do {                                                                          \
    const int io__ = (IN) * 3 + (OUT);                                          \
    switch (io__) {                                                             \
      case DCS_F32 * 3 + DCS_F32: F(float, float); break;                       \
      case DCS_F32 * 3 + DCS_BF16: F(float, __nv_bfloat16); break;              \
      case DCS_F32 * 3 + DCS_F16: F(float, __half); break;                      \
      case DCS_BF16 * 3 + DCS_F32: F(__nv_bfloat16, float); break;              \
      case DCS_BF16 * 3 + DCS_BF16: F(__nv_bfloat16, __nv_bfloat16); break;     \
      case DCS_BF16 * 3 + DCS_F16: F(__nv_bfloat16, __half); break;             \
      case DCS_F16 * 3 + DCS_F32: F(__half, float); break;                      \
      case DCS_F16 * 3 + DCS_BF16: F(__half, __nv_bfloat16); break;             \
      default: F(__half, __half); break;                                        \
    }                                                                           \
  } while (0)

extern "C" int dcs_abi_version(void) { return DCS_ABI_VERSION; }
extern "C" int dcs_zero(void* dst, int64_t bytes, void* stream) {
  DCS_REQUIRE(dst && bytes > 0, "dcs_zero: bad arguments");
  DCS_CUDA(cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}
extern "C" const char* dcs_last_error_string(void) { return err_buf(); }
extern "C" uint64_t dcs_launch_count(void) { return g_launches.load(); }

extern "C" int dcs_cbn_apply(const dcs_cbn_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->affine, "dcs_cbn_apply: null pointer");
  DCS_REQUIRE(p->n_pix > 0 && p->channels > 0, "dcs_cbn_apply: bad shape");
  const int64_t n = p->n_pix * p->channels;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = ew_grid(n, 256);
  DCS_REQUIRE(is_dtype(p->in_dtype) && is_dtype(p->out_dtype), "dcs_cbn_apply: bad dtype");
#define DCS_CBN(TI, TO) cbn_kernel<TI, TO><<<g, 256, 0, s>>>((const TI*)p->x, (TO*)p->y, p->affine, n, p->channels, p->act)
  DCS_DISPATCH_IO(p->in_dtype, p->out_dtype, DCS_CBN);
#undef DCS_CBN
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_mask_combine(const dcs_mask_combine_params* p, void* stream) {
  DCS_REQUIRE(p && p->net_raw && p->noisy_spec && p->clean_spec, "dcs_mask_combine: null pointer");
  DCS_REQUIRE(p->n > 0, "dcs_mask_combine: empty input");
  DCS_REQUIRE(p->combine == DCS_COMBINE_DCS || p->combine == DCS_COMBINE_DC, "dcs_mask_combine: bad combine mode %d", p->combine);
  mask_combine_kernel<<<ew_grid(p->n, 256), 256, 0, (cudaStream_t)stream>>>(*p);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_convert(const void* src, void* dst, int64_t n_floats, int in_dtype, int out_dtype, void* stream) {
  DCS_REQUIRE(src && dst && n_floats > 0 && n_floats % 2 == 0, "dcs_convert: bad arguments");
  const int64_t n2 = n_floats / 2;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = ew_grid(n2, 256);
  DCS_REQUIRE(is_dtype(in_dtype) && is_dtype(out_dtype), "dcs_convert: bad dtype");
#define DCS_CVT(TI, TO) convert_kernel<TI, TO><<<g, 256, 0, s>>>((const TI*)src, (TO*)dst, n2)
  DCS_DISPATCH_IO(in_dtype, out_dtype, DCS_CVT);
#undef DCS_CVT
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_bound_crm(const float* x, float* y, int64_t n, float atan2_eps, int exact_polar, void* stream) {
  DCS_REQUIRE(x && y && n > 0, "dcs_bound_crm: bad arguments");
  bound_crm_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)x, (float2*)y, n, atan2_eps, exact_polar);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_cmul(const float* a, const float* b, float* y, int64_t n, void* stream) {
  DCS_REQUIRE(a && b && y && n > 0, "dcs_cmul: bad arguments");
  cmul_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)a, (const float2*)b, (float2*)y, n);
  DCS_LAUNCHED();
  return 0;
}
// ---- real path (dr / drs) step functions: network_functions.py:286-288 (magnitude, phase = atan2(im, re + eps)) and
//      296-305 / 338-342 (magnitude-mask combine).  mask may be a strided view: element i of image-major (B, F*T) at
//      mask[i * mask_stride].
namespace dcs {
__global__ void mag_phase_kernel(const float2* __restrict__ s, float* __restrict__ mag, float* __restrict__ phase, int64_t n, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = s[i];
    mag[i] = hypotf(v.x, v.y);                  // torch.abs(complex64)
    if (phase) phase[i] = atan2f(v.y, v.x + eps);
  }
}
__global__ void real_mask_combine_kernel(const float* __restrict__ mag, const float* __restrict__ mask, int64_t mask_stride,
                                         float* __restrict__ clean, float* __restrict__ noise, int64_t n, int subtract) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mag[i], pm = m * mask[i * mask_stride];
    if (subtract) { clean[i] = m - pm; if (noise) noise[i] = pm; }   // drs: noise mask, clean = noisy - noise
    else clean[i] = pm;                                              // dr: the mask applies to the speech directly
  }
}
}  // namespace dcs
extern "C" int dcs_mag_phase(const float* spec, float* mag, float* phase, int64_t n, float atan2_eps, void* stream) {
  DCS_REQUIRE(spec && mag && n > 0, "dcs_mag_phase: bad arguments");
  mag_phase_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)spec, mag, phase, n, atan2_eps);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_real_mask_combine(const float* mag, const float* mask, int64_t mask_stride, float* clean_mag, float* noise_mag,
                                     int64_t n, int subtract, void* stream) {
  DCS_REQUIRE(mag && mask && clean_mag && n > 0 && mask_stride > 0, "dcs_real_mask_combine: bad arguments");
  real_mask_combine_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(mag, mask, mask_stride, clean_mag, noise_mag, n, subtract);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_crm(const float* s, const float* y_noisy, float* m, int64_t n, float eps, void* stream) {
  DCS_REQUIRE(s && y_noisy && m && n > 0, "dcs_crm: bad arguments");
  crm_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)s, (const float2*)y_noisy, (float2*)m, n, eps);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_upsample_nearest(const void* x, void* y, int batch, int h, int w, int channels, int up_h, int up_w,
                                    int dtype, void* stream) {
  DCS_REQUIRE(x && y && batch > 0 && h > 0 && w > 0 && channels > 0 && up_h >= 1 && up_w >= 1, "dcs_upsample_nearest: bad arguments");
  const int64_t n = (int64_t)batch * h * up_h * w * up_w * channels;
  cudaStream_t s = (cudaStream_t)stream;
  DCS_REQUIRE(is_dtype(dtype), "dcs_upsample_nearest: bad dtype");
  if (dtype == DCS_BF16) upsample_kernel<__nv_bfloat16><<<ew_grid(n, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, batch, h, w, channels, up_h, up_w);
  else if (dtype == DCS_F16) upsample_kernel<__half><<<ew_grid(n, 256), 256, 0, s>>>((const __half*)x, (__half*)y, batch, h, w, channels, up_h, up_w);
  else upsample_kernel<float><<<ew_grid(n, 256), 256, 0, s>>>((const float*)x, (float*)y, batch, h, w, channels, up_h, up_w);
  DCS_LAUNCHED();
  return 0;
}
