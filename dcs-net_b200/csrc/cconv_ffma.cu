// cconv_ffma.cu — complex convolution as one real implicit GEMM on the fp32 CUDA cores (the <=1e-5 "fp32 mode",
// and the layers whose GEMM shape cannot feed tcgen05: enc0 with 2*Cin = 2).
//
// Replaces apply_complex(conv_r, conv_i) / apply_complex(conv_tran_r, conv_tran_i) / apply_complex(fc_r, fc_i)
// (complexPyTorch 0.3; call sites /root/reference/c_network.py:107-112, 135-147, 202) plus the torch.cat +
// complex_upsample in front of each decoder layer (c_network.py:214-216).  See include/dcsnet.h for geometry.
//
// Tile: 64 pixels x BN outputs per CTA (256 threads, 4 x BN/16 register micro-tile), K-step 16 = one 16-float
// slice of one tap (channels-last => 64 contiguous bytes per pixel), register-prefetched double buffering.
#include <type_traits>
#include "common.cuh"

namespace dcs {

constexpr int kBM = 64, kBK = 16, kConvThreads = 256;

struct ConvGeom {
  int PH, PW;        // phase grid (out_h/up_h, out_w/up_w)
  int64_t m_total;   // B*PH*PW pixels per phase
};

template <typename TIN>
__device__ __forceinline__ float4 load4(const TIN* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_h2<__nv_bfloat16>(r.x), b = unpack_h2<__nv_bfloat16>(r.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_h2<__half>(r.x), b = unpack_h2<__half>(r.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename TIN>
__device__ __forceinline__ float load1(const TIN* p);
template <>
__device__ __forceinline__ float load1<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float load1<__half>(const __half* p) { return __half2float(*p); }

// `vec_ok`: the destination is 4-element aligned (2*cout % 4 == 0); otherwise (odd cout) every other pixel's 4 outputs
// straddle an alignment boundary and a vector store would fault
template <typename TOUT>
__device__ __forceinline__ void store_n(TOUT* p, const float* v, int n, bool vec_ok);
template <>
__device__ __forceinline__ void store_n<float>(float* p, const float* v, int n, bool vec_ok) {
  if (n == 4 && vec_ok) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else for (int i = 0; i < n; ++i) p[i] = v[i];
}
template <typename T>
__device__ __forceinline__ void store_n_h16(T* p, const float* v, int n, bool vec_ok) {
  if (n == 4 && vec_ok) *reinterpret_cast<uint2*>(p) = make_uint2(pack_h2<T>(v[0], v[1]), pack_h2<T>(v[2], v[3]));
  else for (int i = 0; i < n; ++i) p[i] = from_float<T>(v[i]);
}
template <>
__device__ __forceinline__ void store_n<__nv_bfloat16>(__nv_bfloat16* p, const float* v, int n, bool vec_ok) { store_n_h16(p, v, n, vec_ok); }
template <>
__device__ __forceinline__ void store_n<__half>(__half* p, const float* v, int n, bool vec_ok) { store_n_h16(p, v, n, vec_ok); }

// VEC: 2*(c0+c1) and 2*c0 are multiples of 16 -> one K-step lies inside one tap and one source.
template <int BN, bool VEC, typename TIN, typename TOUT>
__global__ void __launch_bounds__(kConvThreads) cconv_ffma_kernel(const dcs_cconv_params p, const ConvGeom g, int n_pad) {
  constexpr int TN = BN / 16;  // outputs per thread along N
  __shared__ __align__(16) float As[2][kBK][kBM + 4];
  __shared__ __align__(16) float Bs[2][kBK][BN];
  const int tid = threadIdx.x;
  const int phase = blockIdx.z;
  const int ph = phase / p.up_w, pw = phase % p.up_w;
  const int64_t m0 = (int64_t)blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int C2 = 2 * (p.c0 + p.c1);
  const int K = p.ntaps * C2;
  const float* __restrict__ W = reinterpret_cast<const float*>(p.weight) + (int64_t)phase * K * n_pad;
  const TIN* __restrict__ s0 = reinterpret_cast<const TIN*>(p.src0);
  const TIN* __restrict__ s1 = reinterpret_cast<const TIN*>(p.src1);

  // A-load role: pixel row lm, 4-float slice lq of the K-step
  const int lm = tid & 63, lq = tid >> 6;
  int64_t pm = m0 + lm;
  const bool row_ok = pm < g.m_total;
  if (!row_ok) pm = 0;
  const int pb = (int)(pm / ((int64_t)g.PH * g.PW));
  const int pj = (int)((pm / g.PW) % g.PH), pi = (int)(pm % g.PW);
  const int sy0 = pj * p.stride_h, sx0 = pi * p.stride_w;
  // B-load role
  const int bk = tid / (BN / 4), bn4 = tid % (BN / 4);
  const bool b_loader = tid < kBK * (BN / 4);

  auto gather_a = [&](int kstep) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = kstep * kBK + 4 * lq;
    if (VEC) {
      const int tap = k / C2, c = k - tap * C2;
      const int y = sy0 + p.dy[phase * p.ntaps + tap], x = sx0 + p.dx[phase * p.ntaps + tap];
      if (row_ok && (unsigned)y < (unsigned)p.in_h && (unsigned)x < (unsigned)p.in_w) {
        const int64_t pix = ((int64_t)pb * p.in_h + y) * p.in_w + x;
        if (c < 2 * p.c0) v = load4<TIN>(s0 + pix * (2 * p.c0) + c);
        else v = load4<TIN>(s1 + pix * (2 * p.c1) + (c - 2 * p.c0));
      }
    } else {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = k + e;
        if (kk < K && row_ok) {
          const int tap = kk / C2, c = kk - tap * C2;
          const int y = sy0 + p.dy[phase * p.ntaps + tap], x = sx0 + p.dx[phase * p.ntaps + tap];
          if ((unsigned)y < (unsigned)p.in_h && (unsigned)x < (unsigned)p.in_w) {
            const int64_t pix = ((int64_t)pb * p.in_h + y) * p.in_w + x;
            t[e] = (c < 2 * p.c0) ? load1<TIN>(s0 + pix * (2 * p.c0) + c) : load1<TIN>(s1 + pix * (2 * p.c1) + (c - 2 * p.c0));
          }
        }
      }
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
    return v;
  };
  auto gather_b = [&](int kstep) -> float4 {
    const int k = kstep * kBK + bk;
    if (b_loader && k < K) return __ldg(reinterpret_cast<const float4*>(W + (int64_t)k * n_pad + n0 + 4 * bn4));
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf, float4 a, float4 b) {
    As[buf][4 * lq + 0][lm] = a.x; As[buf][4 * lq + 1][lm] = a.y;
    As[buf][4 * lq + 2][lm] = a.z; As[buf][4 * lq + 3][lm] = a.w;
    if (b_loader) *reinterpret_cast<float4*>(&Bs[buf][bk][4 * bn4]) = b;
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nsteps = (K + kBK - 1) / kBK;
  float4 ra = gather_a(0), rb = gather_b(0);
  stash(0, ra, rb);
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) { ra = gather_a(s + 1); rb = gather_b(s + 1); }
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
      float b[TN];
      if (TN == 4) { const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]); b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w; }
      else if (TN == 2) { const float2 t = *reinterpret_cast<const float2*>(&Bs[buf][k][2 * tx]); b[0] = t.x; b[1] = t.y; }
      else b[0] = Bs[buf][k][tx];
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], b[j], acc[i][j]);
    }
    if (s + 1 < nsteps) {
      stash(buf ^ 1, ra, rb);
      __syncthreads();
    }
  }

  // epilogue: + bias, activation, scatter to the output pixel of this phase
  const int N = 2 * p.cout;
  const int nb = n0 + TN * tx;
  TOUT* __restrict__ dst = reinterpret_cast<TOUT*>(p.dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + 4 * ty + i;
    if (m >= g.m_total) continue;
    const int b = (int)(m / ((int64_t)g.PH * g.PW));
    const int j = (int)((m / g.PW) % g.PH), ii = (int)(m % g.PW);
    const int oy = j * p.up_h + ph, ox = ii * p.up_w + pw;
    float v[TN];
    int nvalid = 0;
#pragma unroll
    for (int jn = 0; jn < TN; ++jn) {
      const int n = nb + jn;
      if (n < N) { v[jn] = act_apply(acc[i][jn] + (p.bias ? __ldg(p.bias + n) : 0.f), p.act); ++nvalid; }
    }
    if (nvalid) store_n<TOUT>(dst + (((int64_t)b * p.out_h + oy) * p.out_w + ox) * N + nb, v, nvalid, (N & 3) == 0);
  }
}

template <int BN, bool VEC, typename TIN>
static int launch_out(const dcs_cconv_params& p, const ConvGeom& g, int n_pad, dim3 grid, cudaStream_t s) {
  // a 16-bit input only pairs with fp32 or the same 16-bit type on the output side
  if constexpr (!std::is_same<TIN, __half>::value) {
    if (p.out_dtype == DCS_BF16) { cconv_ffma_kernel<BN, VEC, TIN, __nv_bfloat16><<<grid, kConvThreads, 0, s>>>(p, g, n_pad); return 0; }
  }
  if constexpr (!std::is_same<TIN, __nv_bfloat16>::value) {
    if (p.out_dtype == DCS_F16) { cconv_ffma_kernel<BN, VEC, TIN, __half><<<grid, kConvThreads, 0, s>>>(p, g, n_pad); return 0; }
  }
  cconv_ffma_kernel<BN, VEC, TIN, float><<<grid, kConvThreads, 0, s>>>(p, g, n_pad);
  return 0;
}
template <int BN, bool VEC>
static int launch_in(const dcs_cconv_params& p, const ConvGeom& g, int n_pad, dim3 grid, cudaStream_t s) {
  if (p.in_dtype == DCS_BF16) return launch_out<BN, VEC, __nv_bfloat16>(p, g, n_pad, grid, s);
  if (p.in_dtype == DCS_F16) return launch_out<BN, VEC, __half>(p, g, n_pad, grid, s);
  return launch_out<BN, VEC, float>(p, g, n_pad, grid, s);
}

int validate_conv(const dcs_cconv_params* p, const char* who) {
  DCS_REQUIRE(p && p->src0 && p->weight && p->dst, "%s: null pointer", who);
  DCS_REQUIRE(p->c0 > 0 && p->c1 >= 0 && (p->c1 == 0 || p->src1), "%s: bad source channels (%d,%d)", who, p->c0, p->c1);
  DCS_REQUIRE(p->batch > 0 && p->in_h > 0 && p->in_w > 0 && p->cout > 0, "%s: bad shape", who);
  DCS_REQUIRE(p->up_h >= 1 && p->up_w >= 1 && p->up_h <= 8 && p->up_w <= 8, "%s: up factors must be in [1, 8]", who);
  DCS_REQUIRE(p->stride_h >= 1 && p->stride_w >= 1, "%s: bad stride", who);
  DCS_REQUIRE(p->out_h % p->up_h == 0 && p->out_w % p->up_w == 0, "%s: out dims not divisible by up factors", who);
  DCS_REQUIRE(p->ntaps >= 1 && p->ntaps * p->up_h * p->up_w <= DCS_MAX_TAPS, "%s: too many taps (%d x %d phases)", who,
              p->ntaps, p->up_h * p->up_w);
  DCS_REQUIRE(is_dtype(p->in_dtype), "%s: bad in_dtype", who);
  DCS_REQUIRE(is_dtype(p->out_dtype), "%s: bad out_dtype", who);
  DCS_REQUIRE(!is_h16(p->in_dtype) || !is_h16(p->out_dtype) || p->in_dtype == p->out_dtype, "%s: mixed 16-bit storage types", who);
  return 0;
}

}  // namespace dcs

using namespace dcs;

extern "C" int dcs_cconv2d_fwd(const dcs_cconv_params* p, void* stream) {
  if (int e = validate_conv(p, "dcs_cconv2d_fwd")) return e;
  DCS_REQUIRE(p->up_h <= 2 && p->up_w <= 2, "dcs_cconv2d_fwd: up factors must be 1 or 2");
  DCS_REQUIRE(!p->pool_sums, "dcs_cconv2d_fwd: fused pooling is only provided by the tcgen05 path; use dcs_chan_pool");
  ConvGeom g;
  g.PH = p->out_h / p->up_h;
  g.PW = p->out_w / p->up_w;
  g.m_total = (int64_t)p->batch * g.PH * g.PW;
  const int N = 2 * p->cout;
  const int n_pad = (N + 15) / 16 * 16;
  const int C2 = 2 * (p->c0 + p->c1);
  const bool vec = (C2 % 16 == 0) && ((2 * p->c0) % 16 == 0);
  const int BN = n_pad >= 64 ? 64 : (n_pad >= 32 ? 32 : 16);
  DCS_REQUIRE(n_pad % BN == 0, "dcs_cconv2d_fwd: unsupported 2*cout=%d", N);
  const int64_t mt = (g.m_total + kBM - 1) / kBM;
  DCS_REQUIRE(mt < (1ll << 31), "dcs_cconv2d_fwd: too many pixels");
  dim3 grid((unsigned)mt, n_pad / BN, p->up_h * p->up_w);
  cudaStream_t s = (cudaStream_t)stream;
  if (BN == 64) { if (vec) launch_in<64, true>(*p, g, n_pad, grid, s); else launch_in<64, false>(*p, g, n_pad, grid, s); }
  else if (BN == 32) { if (vec) launch_in<32, true>(*p, g, n_pad, grid, s); else launch_in<32, false>(*p, g, n_pad, grid, s); }
  else { if (vec) launch_in<16, true>(*p, g, n_pad, grid, s); else launch_in<16, false>(*p, g, n_pad, grid, s); }
  DCS_LAUNCHED();
  return 0;
}
