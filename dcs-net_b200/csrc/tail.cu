// tail.cu — the last decoder layer fused with the mask tail (CUDA cores; N = 2 cannot feed a tensor core):
//
//   decoder[6] = ComplexConvTranspose2d(16 -> 1, k3 s1 p1) on cat(d, skip_sa) up-sampled (2,2)
//                                                      (/root/reference/c_network.py:214-216, 134-140)
//   net_out    = bound_cRM(squeeze(.))                                         (c_network.py:224-226)
//   mask       = bound_cRM(net_out)                                            (network_functions.py:394)
//   N = Y (.) mask,  S = Y - N   |  dc: S = Y (.) mask                         (network_functions.py:396-397, 434)
//
// One pass: reads d, skip (B,H,W,8) and Y, writes S (and optionally N / mask / net_out); decoder[6]'s raw output
// never exists in HBM.  Sub-pixel form: source pixel offsets {-1,0,1}^2 feed the 4 output phases with pre-summed
// 2x2 taps.  A thread owns 4 adjacent source pixels (= a 2 x 8 output patch, 64 contiguous bytes per output row);
// the input tile is transposed into channel planes in shared memory so the inner loop is LDS.128 + FFMA.
//
// Algorithmic bytes per source pixel: 2 x 8 ch x sizeof(act) in, per output pixel 8 (Y) + 8 (S).
#include "common.cuh"

namespace dcs {

constexpr int kTlRows = 4, kTlCols = 128, kTlThreads = 128;
constexpr int kTlPitch = kTlCols + 8;  // tile col c lives at index c + 4 (16-byte aligned quads), halo at 3 and 132
constexpr int kTlCi = 16;

// shared-memory element: one complex activation
template <typename T> struct SmemC;
template <> struct SmemC<__nv_bfloat16> {
  using type = uint32_t;
  static __device__ __forceinline__ float2 get(uint32_t v) { return unpack_h2<__nv_bfloat16>(v); }
};
template <> struct SmemC<__half> {
  using type = uint32_t;
  static __device__ __forceinline__ float2 get(uint32_t v) { return unpack_h2<__half>(v); }
};
template <> struct SmemC<float> {
  using type = float2;
  static __device__ __forceinline__ float2 get(float2 v) { return v; }
};

template <typename T>
__global__ void __launch_bounds__(kTlThreads) dec6_tail_kernel(const dcs_dec6_tail_params p) {
  using E = typename SmemC<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* wsm = reinterpret_cast<float4*>(smem_raw);                         // [ci][phase][tap] (M00 M10 M01 M11)
  E* tile = reinterpret_cast<E*>(smem_raw + kTlCi * 16 * sizeof(float4));    // [ci][row 0..5][kTlPitch]
  const int tid = threadIdx.x;
  const int b = blockIdx.z, r0 = blockIdx.y * kTlRows, c0 = blockIdx.x * kTlCols;
  const int H = p.h, W = p.w;

  for (int i = tid; i < kTlCi * 16; i += kTlThreads) {
    const int ci = i >> 4, pt = i & 15;  // global layout [phase][tap][ci][4]
    wsm[i] = __ldg(reinterpret_cast<const float4*>(p.weight) + pt * kTlCi + ci);
  }
  // ---- load + transpose the (6 x 130) x 2 sources x 8 channels halo tile into channel planes (zero outside)
  const T* srcs[2] = {reinterpret_cast<const T*>(p.d), reinterpret_cast<const T*>(p.skip)};
  for (int it = tid; it < 2 * (kTlRows + 2) * (kTlCols + 2); it += kTlThreads) {
    const int col = it % (kTlCols + 2);
    const int rs = it / (kTlCols + 2);
    const int row = rs % (kTlRows + 2), s = rs / (kTlRows + 2);
    const int y = r0 + row - 1, x = c0 + col - 1;
    E v[8];
    const bool in = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
    const T* g = srcs[s] + (((int64_t)b * H + (in ? y : 0)) * W + (in ? x : 0)) * 16;
    if constexpr (sizeof(E) == 4) {  // bf16: a pixel's 8 complex channels = 2 x 16 bytes, kept packed
      uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
      if (in) { q0 = __ldg(reinterpret_cast<const uint4*>(g)); q1 = __ldg(reinterpret_cast<const uint4*>(g) + 1); }
      v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) q = __ldg(reinterpret_cast<const float4*>(g) + k);
        v[2 * k] = make_float2(q.x, q.y); v[2 * k + 1] = make_float2(q.z, q.w);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[((s * 8 + k) * (kTlRows + 2) + row) * kTlPitch + col + 3] = v[k];
  }
  __syncthreads();

  const int r = tid >> 5, cx = tid & 31;  // source row r0 + r, source cols c0 + 4*cx + {0..3}
  float2 acc[4][4];                       // [pixel][phase]
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[q][f] = make_float2(0.f, 0.f);

#pragma unroll 1
  for (int ci = 0; ci < kTlCi; ++ci) {
    float4 w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = wsm[ci * 16 + i];
    const E* plane = tile + (ci * (kTlRows + 2) + r) * kTlPitch + 4 * cx + 3;
#pragma unroll
    for (int sy = 0; sy < 3; ++sy) {  // source row offset sy - 1
      const E* rowp = plane + sy * kTlPitch;
      float2 xin[6];
      xin[0] = SmemC<T>::get(rowp[0]);
      if constexpr (sizeof(E) == 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(rowp + 1);
        xin[1] = SmemC<T>::get(q.x); xin[2] = SmemC<T>::get(q.y); xin[3] = SmemC<T>::get(q.z); xin[4] = SmemC<T>::get(q.w);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) xin[1 + q] = SmemC<T>::get(rowp[1 + q]);
      }
      xin[5] = SmemC<T>::get(rowp[5]);
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const int ty = sy - ph;  // tap row index: dy = ph - 1 + ty  ==  sy - 1
        if (ty < 0 || ty > 1) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int sx = 0; sx < 3; ++sx) {  // source col offset sx - 1 relative to pixel q -> xin[q + sx]
            const float2 x = xin[q + sx];
            const float2 xr = make_float2(x.x, x.x), xi = make_float2(x.y, x.y);
#pragma unroll
            for (int pw = 0; pw < 2; ++pw) {
              const int tx = sx - pw;
              if (tx < 0 || tx > 1) continue;
              const float4 m = w[(ph * 2 + pw) * 4 + ty * 2 + tx];  // (M00, M10, M01, M11): column pairs of the 2x2 block
              float2& a = acc[q][ph * 2 + pw];
              ffma2(a, make_float2(m.z, m.w), xi);   // (re, im) += (M01, M11) * x.im
              ffma2(a, make_float2(m.x, m.y), xr);   // (re, im) += (M00, M10) * x.re
            }
          }
        }
      }
    }
  }

  // ---- epilogue: bias, bound_cRM twice, complex product with Y, subtraction; 2 output rows x 8 contiguous columns
  const int j = r0 + r;
  if (j >= H) return;
  const int OW = 2 * W;
  const bool exact = p.exact_polar != 0;
  const float2* Y = reinterpret_cast<const float2*>(p.noisy_spec);
  float2* o_clean = reinterpret_cast<float2*>(p.clean_spec);
  float2* o_noise = reinterpret_cast<float2*>(p.noise_spec);
  float2* o_mask = reinterpret_cast<float2*>(p.mask);
  float2* o_net = reinterpret_cast<float2*>(p.net_out);
  float2* o_raw = reinterpret_cast<float2*>(p.net_raw);
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
    const int64_t rowbase = ((int64_t)b * 2 * H + 2 * j + ph) * OW;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = c0 + 4 * cx + q;
      if (i >= W) continue;
      const int64_t o = rowbase + 2 * i;  // two adjacent output columns (pw = 0, 1): one 16-byte access per array
      const float4 y2 = __ldg(reinterpret_cast<const float4*>(Y + o));
      float2 raw[2], m1[2], m2[2], cl[2], ns[2];
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        raw[pw] = make_float2(acc[q][ph * 2 + pw].x + p.bias_re, acc[q][ph * 2 + pw].y + p.bias_im);
        m1[pw] = bound_crm_dev(raw[pw], p.atan2_eps, exact);
        m2[pw] = bound_crm_dev(m1[pw], p.atan2_eps, exact);
        const float2 y = pw ? make_float2(y2.z, y2.w) : make_float2(y2.x, y2.y);
        const float2 pr = cmul(y, m2[pw]);
        if (p.combine == DCS_COMBINE_DCS) { ns[pw] = pr; cl[pw] = make_float2(y.x - pr.x, y.y - pr.y); }
        else { ns[pw] = pr; cl[pw] = pr; }
      }
      *reinterpret_cast<float4*>(o_clean + o) = make_float4(cl[0].x, cl[0].y, cl[1].x, cl[1].y);
      if (o_noise && p.combine == DCS_COMBINE_DCS) *reinterpret_cast<float4*>(o_noise + o) = make_float4(ns[0].x, ns[0].y, ns[1].x, ns[1].y);
      if (o_mask) *reinterpret_cast<float4*>(o_mask + o) = make_float4(m2[0].x, m2[0].y, m2[1].x, m2[1].y);
      if (o_net) *reinterpret_cast<float4*>(o_net + o) = make_float4(m1[0].x, m1[0].y, m1[1].x, m1[1].y);
      if (o_raw) *reinterpret_cast<float4*>(o_raw + o) = make_float4(raw[0].x, raw[0].y, raw[1].x, raw[1].y);
    }
  }
}

}  // namespace dcs

using namespace dcs;

extern "C" int dcs_dec6_tail_fwd(const dcs_dec6_tail_params* p, void* stream) {
  DCS_REQUIRE(p && p->d && p->skip && p->weight && p->noisy_spec && p->clean_spec, "dcs_dec6_tail_fwd: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->batch <= 65535 && p->h > 0 && p->w > 0, "dcs_dec6_tail_fwd: bad shape");
  DCS_REQUIRE(p->combine == DCS_COMBINE_DCS || p->combine == DCS_COMBINE_DC, "dcs_dec6_tail_fwd: bad combine mode");
  DCS_REQUIRE(is_dtype(p->in_dtype), "dcs_dec6_tail_fwd: bad dtype");
  dim3 grid((p->w + kTlCols - 1) / kTlCols, (p->h + kTlRows - 1) / kTlRows, p->batch);
  const size_t elem = is_h16(p->in_dtype) ? 4 : 8;
  const size_t smem = kTlCi * 16 * sizeof(float4) + (size_t)kTlCi * (kTlRows + 2) * kTlPitch * elem;
  cudaStream_t s = (cudaStream_t)stream;
  if (p->in_dtype == DCS_BF16) {
    DCS_CUDA(cudaFuncSetAttribute(dec6_tail_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec6_tail_kernel<__nv_bfloat16><<<grid, kTlThreads, smem, s>>>(*p);
  } else if (p->in_dtype == DCS_F16) {
    DCS_CUDA(cudaFuncSetAttribute(dec6_tail_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec6_tail_kernel<__half><<<grid, kTlThreads, smem, s>>>(*p);
  } else {
    DCS_CUDA(cudaFuncSetAttribute(dec6_tail_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec6_tail_kernel<float><<<grid, kTlThreads, smem, s>>>(*p);
  }
  DCS_LAUNCHED();
  return 0;
}
