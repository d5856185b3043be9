// cconv_first.cu — encoder[0] fused with initial_batchnorm (CUDA cores).
//
//   e0 = initial_batchnorm(x.view(B,1,F,T))                                   (/root/reference/c_network.py:190)
//   e1 = ComplexReLU(ComplexBatchNorm2d(ComplexConv2d(1 -> 8, k7, s(2,2), p3)(e0)))   (c_network.py:107-114)
//
// 2*Cin = 2 gives a K of 98 made of 4-byte pieces: no TMA box / UMMA K-slice fits, so this layer is a direct
// convolution on the fp32 pipes.  The folded 2x2 affine of initial_batchnorm is applied while the spectrogram tile is
// staged (zero padding stays zero AFTER the affine, as in the reference), so the BN'd input never exists in HBM
// and the bf16 mode does not round the network input.
//
// CTA: 16 x 64 output pixels, 256 threads; lane <-> output column, each thread owns 4 output rows x 8 complex
// channels (64 fp32 accumulators).  The input tile is split by column parity so that a stride-2 tap reads
// consecutive shared-memory words across a warp; weights are broadcast LDS.128.
// Algorithmic bytes per output pixel: 4 input pixels x 8 B read + 8 ch x sizeof(act) x 2 written.
#include "common.cuh"

namespace dcs {

constexpr int kF_TH = 16, kF_TW = 64, kF_K = 7, kF_Threads = 256;
constexpr int kF_IH = 2 * kF_TH + 5;            // 37 input rows
constexpr int kF_HW = kF_TW + 3;                // 67 columns per parity plane
constexpr int kF_N = 16;

struct FirstSmem {
  float4 w[49 * 2 * 4];                  // [tap][ri][16 n]
  float2 ev[kF_IH][kF_HW + 1];           // even tile columns
  float2 od[kF_IH][kF_HW + 1];           // odd tile columns
};

template <typename TOUT>
__global__ void __launch_bounds__(kF_Threads, 2) enc0_kernel(const dcs_enc0_params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FirstSmem& sm = *reinterpret_cast<FirstSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.z, oy0 = blockIdx.y * kF_TH, ox0 = blockIdx.x * kF_TW;
  const int H = p.h, W = p.w, OH = H / 2, OW = W / 2;
  for (int i = tid; i < 49 * 2 * 4; i += kF_Threads) sm.w[i] = __ldg(reinterpret_cast<const float4*>(p.weight) + i);
  float A00 = 1.f, A01 = 0.f, A10 = 0.f, A11 = 1.f, c0 = 0.f, c1 = 0.f;
  if (p.bn_affine) { A00 = p.bn_affine[0]; A01 = p.bn_affine[1]; A10 = p.bn_affine[2]; A11 = p.bn_affine[3]; c0 = p.bn_affine[4]; c1 = p.bn_affine[5]; }
  const float2* Y = reinterpret_cast<const float2*>(p.spec) + (int64_t)b * H * W;
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
  for (int i = tid; i < kF_IH * (2 * kF_HW); i += kF_Threads) {
    const int r = i / (2 * kF_HW), c = i - r * (2 * kF_HW);  // tile col c = 2*ox_l + kx in [0, 134)
    const int y = iy0 + r, x = ix0 + c;
    float2 v = make_float2(0.f, 0.f);
    if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
      const float2 s = __ldg(Y + (int64_t)y * W + x);
      v = make_float2(A00 * s.x + A01 * s.y + c0, A10 * s.x + A11 * s.y + c1);
    }
    if (c & 1) sm.od[r][c >> 1] = v; else sm.ev[r][c >> 1] = v;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int lx = (warp & 1) * 32 + lane;       // local output column
  const int ly = (warp >> 1) * 4;              // first of 4 local output rows
  float2 acc[4][kF_N / 2];  // (n, n+1) pairs: packed fp32x2 FMA halves the FMA instruction count
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int n = 0; n < kF_N / 2; ++n) acc[q][n] = make_float2(0.f, 0.f);

#pragma unroll 1
  for (int ky = 0; ky < kF_K; ++ky) {
#pragma unroll
    for (int kx = 0; kx < kF_K; ++kx) {
      float2 wr[kF_N / 2], wi[kF_N / 2];
      const float4* wp = sm.w + (ky * kF_K + kx) * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = wp[j], c = wp[4 + j];
        wr[2 * j] = make_float2(a.x, a.y); wr[2 * j + 1] = make_float2(a.z, a.w);
        wi[2 * j] = make_float2(c.x, c.y); wi[2 * j + 1] = make_float2(c.z, c.w);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = 2 * (ly + q) + ky;
        const float2 x = (kx & 1) ? sm.od[r][lx + (kx >> 1)] : sm.ev[r][lx + (kx >> 1)];
        const float2 xr = make_float2(x.x, x.x), xi = make_float2(x.y, x.y);
#pragma unroll
        for (int n = 0; n < kF_N / 2; ++n) { ffma2(acc[q][n], wi[n], xi); ffma2(acc[q][n], wr[n], xr); }
      }
    }
  }

  const int ox = ox0 + lx;
  if (ox >= OW) return;
  TOUT* dst = reinterpret_cast<TOUT*>(p.dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oy = oy0 + ly + q;
    if (oy >= OH) continue;
    float v[kF_N];
#pragma unroll
    for (int n = 0; n < kF_N / 2; ++n) {
      v[2 * n] = fmaxf(acc[q][n].x + __ldg(p.bias + 2 * n), 0.f);
      v[2 * n + 1] = fmaxf(acc[q][n].y + __ldg(p.bias + 2 * n + 1), 0.f);
    }
    TOUT* o = dst + (((int64_t)b * OH + oy) * OW + ox) * kF_N;
    if constexpr (sizeof(TOUT) == 2) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_h2<TOUT>(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(o + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
}

}  // namespace dcs

using namespace dcs;

extern "C" int dcs_enc0_fwd(const dcs_enc0_params* p, void* stream) {
  DCS_REQUIRE(p && p->spec && p->weight && p->bias && p->dst, "dcs_enc0_fwd: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->batch <= 65535 && p->h > 0 && p->w > 0 && p->h % 2 == 0 && p->w % 2 == 0, "dcs_enc0_fwd: bad shape");
  DCS_REQUIRE(is_dtype(p->out_dtype), "dcs_enc0_fwd: bad out_dtype");
  dim3 grid((p->w / 2 + kF_TW - 1) / kF_TW, (p->h / 2 + kF_TH - 1) / kF_TH, p->batch);
  const size_t smem = sizeof(FirstSmem);
  cudaStream_t s = (cudaStream_t)stream;
  if (p->out_dtype == DCS_BF16) {
    DCS_CUDA(cudaFuncSetAttribute(enc0_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enc0_kernel<__nv_bfloat16><<<grid, kF_Threads, smem, s>>>(*p);
  } else if (p->out_dtype == DCS_F16) {
    DCS_CUDA(cudaFuncSetAttribute(enc0_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enc0_kernel<__half><<<grid, kF_Threads, smem, s>>>(*p);
  } else {
    DCS_CUDA(cudaFuncSetAttribute(enc0_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enc0_kernel<float><<<grid, kF_Threads, smem, s>>>(*p);
  }
  DCS_LAUNCHED();
  return 0;
}
