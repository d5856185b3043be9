// wgrad_tc.cu — weight gradient of a complex convolution on the 5th-generation tensor cores (training step):
//
//   dWp[n][tap][k] = sum over output pixels (b, j, i) of  dY[b, j, i][n] * X[b, j*sh + dy(tap), i*sw + dx(tap)][k]
//
// in the packed real formulation of the forward kernels (n = 2*cout + re/im, k = 2*cin + re/im; oracle/train_oracle.py
// cconv2d_backward: the four blocks fold into dw_r = dWp[re,re] + dWp[im,im], dw_i = dWp[im,re] - dWp[re,im]).  Replaces the
// autograd backward of apply_complex(conv_r, conv_i) (complexPyTorch 0.3) at /root/reference/c_network.py:107-112.
//
// GEMM view: M = 2*Cout (tiles of 128 rows), N = 2*Cin (<= 256), K = pixels.  Both operands are the channels-last activations
// exactly as they lie in HBM, [pixel][channel]: the GEMM's K (pixels) is the SLOW dimension, so both are MN-major UMMA operands
// (instruction-descriptor bits 15 / 16), staged by TMA as 64-channel x 64-pixel boxes with SWIZZLE_128B:
//   canonical MN-major SW128 layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) elements — a 64-channel x 8-pixel atom is 8 rows of
//   128 bytes; LBO = distance between 64-channel boxes, SBO = 1024 bytes between 8-pixel groups; one tcgen05.mma covers
//   K = 16 pixels = two groups, the next one starts 2048 bytes further.
// Zero padding, image borders and ragged rows come from TMA out-of-bounds fill (a zero on either side contributes nothing),
// the conv stride from the tensor map's element strides.
// Work split: CTA = (tap, M tile, K split); fp32 accumulator in TMEM; partial tiles go to a workspace and a second kernel adds
// the K splits in a fixed order (deterministic) and folds the four real blocks into (dw_r, dw_i).
#include <cuda.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dcs {

constexpr int kWgPix = 64;                   // pixels per pipeline stage (K of the GEMM per stage)
constexpr int kWgBoxBytes = kWgPix * 128;    // one TMA box: 64 pixels x 64 channels x 2 bytes
constexpr int kWgStages = 4;
constexpr int kWgThreads = 192;              // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct WgArgs {
  int n_taps, m_tiles, ksplits, steps_total, steps_per_split;
  int OH, OW, wblocks;                       // output grid; 64-pixel blocks per output row
  int sh, sw, nb;                            // conv stride; 64-channel boxes of X (2*Cin / 64)
  int N;                                     // 2*Cin
  int f16;
  int8_t dy[DCS_MAX_TAPS], dx[DCS_MAX_TAPS];
  float* partial;                            // [ksplit][tap][m_tile][128][N]
};

struct __align__(8) WgBars {
  uint64_t full[kWgStages], empty[kWgStages], acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const WgArgs a) {
  extern __shared__ __align__(1024) unsigned char wg_smem[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)wg_smem + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = (uint32_t)(2 + a.nb) * kWgBoxBytes;
  WgBars* bars = reinterpret_cast<WgBars*>(base + (size_t)kWgStages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)a.N) tmem_cols <<= 1;
  // unit -> (ksplit, tap, m tile)
  int u = blockIdx.x;
  const int mt = u % a.m_tiles; u /= a.m_tiles;
  const int tap = u % a.n_taps;
  const int ks = u / a.n_taps;
  const int s0 = ks * a.steps_per_split, s1 = min(a.steps_total, s0 + a.steps_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    uint32_t stage = 0, par = 0;
    const int dyv = a.dy[tap], dxv = a.dx[tap];
    for (int s = s0; s < s1; ++s) {
      const int ib = s % a.wblocks, r = s / a.wblocks;
      const int j = r % a.OH, b = r / a.OH;
      const int i0 = ib * kWgPix;
      mbar_wait(smem_u32(&bars->empty[stage]), par ^ 1);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bars->full[stage]);
        const uint32_t dst = smem_u32(base) + stage * stage_bytes;
        mbar_expect_tx(full, stage_bytes);
        for (int h = 0; h < 2; ++h) tma_load_4d(dst + h * kWgBoxBytes, &tmY, full, mt * 128 + h * 64, i0, j, b);
        for (int h = 0; h < a.nb; ++h)
          tma_load_4d(dst + (2 + h) * kWgBoxBytes, &tmX, full, h * 64, i0 * a.sw + dxv, j * a.sh + dyv, b);
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; par ^= 1; }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t fmt = a.f16 ? 0u : 1u;
    // D = F32, A / B = F16 | BF16, both MN-major (bits 15, 16), N >> 3 @ 17, M >> 4 @ 24
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t stage = 0, par = 0;
    for (int s = s0; s < s1; ++s) {
      mbar_wait(smem_u32(&bars->full[stage]), par);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(base) + stage * stage_bytes, sb = sa + 2 * kWgBoxBytes;
#pragma unroll
        for (int kk = 0; kk < kWgPix / 16; ++kk) {
          const uint64_t ad = umma_desc_mn(sa + kk * 2048, kWgBoxBytes, 1024);
          const uint64_t bd = umma_desc_mn(sb + kk * 2048, kWgBoxBytes, 1024);
          tc_mma_bf16(tmem_base, ad, bd, idesc, (s != s0 || kk != 0) ? 1u : 0u);
        }
        tc_commit(smem_u32(&bars->empty[stage]));
        if (s == s1 - 1) tc_commit(smem_u32(&bars->acc_full));
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; par ^= 1; }
    }
  } else {
    // ===================================================================== epilogue: TMEM -> partial[ks][tap][mt][row][N]
    const int quad = warp & 3;
    float* out = a.partial + ((((int64_t)ks * a.n_taps + tap) * a.m_tiles + mt) * 128 + quad * 32 + lane) * a.N;
    if (s1 > s0) {
      mbar_wait(smem_u32(&bars->acc_full), 0);
      tc_fence_after();
      for (int c = 0; c < a.N; c += 16) {
        uint32_t rg[16];
        tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, rg);
        tc_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; q += 4)
          *reinterpret_cast<float4*>(out + c + q) = make_float4(__uint_as_float(rg[q]), __uint_as_float(rg[q + 1]), __uint_as_float(rg[q + 2]), __uint_as_float(rg[q + 3]));
      }
    } else {
      for (int c = 0; c < a.N; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// K-split reduction in a fixed order + fold of the four real blocks:  dw_r[co][ci][tap] = dWp[(co,re)][(ci,re)] + dWp[(co,im)][(ci,im)],
// dw_i[co][ci][tap] = dWp[(co,im)][(ci,re)] - dWp[(co,re)][(ci,im)]   (reference weight layout (Cout, Cin, kh, kw), tap = ky*kw + kx)
__global__ void wgrad_fold_kernel(const float* __restrict__ partial, int ksplits, int n_taps, int cout, int cin, float* __restrict__ dw_r,
                                  float* __restrict__ dw_i) {
  const int n_elem = cout * cin * n_taps;
  const int M = 2 * cout, N = 2 * cin;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += gridDim.x * blockDim.x) {
    const int tap = e % n_taps, ci = (e / n_taps) % cin, co = e / (n_taps * cin);
    float rr = 0.f, ii = 0.f, ir = 0.f, ri = 0.f;
    for (int k = 0; k < ksplits; ++k) {
      const float* p = partial + ((int64_t)k * n_taps + tap) * M * N;
      rr += p[(int64_t)(2 * co) * N + 2 * ci]; ri += p[(int64_t)(2 * co) * N + 2 * ci + 1];
      ir += p[(int64_t)(2 * co + 1) * N + 2 * ci]; ii += p[(int64_t)(2 * co + 1) * N + 2 * ci + 1];
    }
    dw_r[e] = rr + ii;
    dw_i[e] = ir - ri;
  }
}

// K-split reduction in a fixed order into the generic layout of dcs_wgrad: dwp[tap][k][n] = sum_pixels x[k] dy[n]  (partial rows = dy
// channel n, columns = x channel k, padded to Mp x Np)
__global__ void wgrad_reduce_generic_kernel(const float* __restrict__ partial, int ksplits, int n_taps, int Mp, int Np, int k2, int n2,
                                            float* __restrict__ dwp, int64_t tap_stride) {
  const int n_elem = n_taps * k2 * n2;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += gridDim.x * blockDim.x) {
    const int n = e % n2, k = (e / n2) % k2, tap = e / (n2 * k2);
    float s = 0.f;
    for (int z = 0; z < ksplits; ++z) s += partial[(((int64_t)z * n_taps + tap) * Mp + n) * Np + k];
    dwp[(int64_t)tap * tap_stride + (int64_t)k * n2 + n] = s;
  }
}

typedef CUresult (*WgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WgEncodeFn wg_encode_fn() {
  static WgEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (WgEncodeFn)p;
  }
  return fn;
}
static int wg_map(CUtensorMap* m, const void* ptr, int f16, int C2, int W, int H, int B, int stride_w, int pitch = 0) {
  WgEncodeFn fn = wg_encode_fn();
  DCS_REQUIRE(fn, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  if (pitch <= 0) pitch = C2;                 // channels per pixel in memory (>= C2 when the operand is a channel slice)
  cuuint64_t dims[4] = {(cuuint64_t)C2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(kWgPix * stride_w), 1, 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride_w, 1, 1};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad) failed with CUresult %d (C2=%d W=%d H=%d B=%d)", (int)r, C2, W, H, B);
  return 0;
}

static int wg_split(const dcs_cwgrad_params* p, int* ksplits, int* steps_total, int* per) {
  const int wblocks = (p->out_w + kWgPix - 1) / kWgPix;
  *steps_total = p->batch * p->out_h * wblocks;
  const int units = p->ntaps * ((2 * p->cout + 127) / 128);
  int ks = std::max(1, (2 * num_sms() + units - 1) / units);
  ks = std::min(ks, std::max(1, *steps_total / 4));
  *per = (*steps_total + ks - 1) / ks;
  *ksplits = (*steps_total + *per - 1) / *per;
  return 0;
}

}  // namespace dcs

using namespace dcs;

extern "C" int64_t dcs_cwgrad_workspace_bytes(const dcs_cwgrad_params* p) {
  if (!p || p->batch <= 0 || p->cout <= 0 || p->cin <= 0) return -1;
  int ks, st, per;
  wg_split(p, &ks, &st, &per);
  return (int64_t)ks * p->ntaps * (2 * p->cout) * (2 * p->cin) * (int64_t)sizeof(float);
}

extern "C" int dcs_cwgrad_tc(const dcs_cwgrad_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->dy && p->dw_r && p->dw_i && p->workspace, "dcs_cwgrad_tc: null pointer");
  DCS_REQUIRE(is_h16(p->dtype), "dcs_cwgrad_tc: x / dy must be DCS_F16 or DCS_BF16 storage");
  DCS_REQUIRE(p->batch > 0 && p->in_h > 0 && p->in_w > 0 && p->out_h > 0 && p->out_w > 0, "dcs_cwgrad_tc: bad shape");
  DCS_REQUIRE((2 * p->cout) % 128 == 0 && (2 * p->cin) % 64 == 0 && 2 * p->cin <= 256, "dcs_cwgrad_tc: needs 2*cout %% 128 == 0 and 2*cin in {64, 128, 192, 256} (got %d, %d)",
              2 * p->cout, 2 * p->cin);
  DCS_REQUIRE(p->ntaps >= 1 && p->ntaps <= DCS_MAX_TAPS && (p->stride_w == 1 || p->stride_w == 2) && p->stride_h >= 1, "dcs_cwgrad_tc: bad taps / stride");
  DCS_REQUIRE(((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->dy % 16 == 0), "dcs_cwgrad_tc: pointers must be 16-byte aligned");
  DCS_REQUIRE(p->workspace_bytes >= dcs_cwgrad_workspace_bytes(p), "dcs_cwgrad_tc: workspace too small");
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.n_taps = p->ntaps; a.m_tiles = (2 * p->cout) / 128;
  wg_split(p, &a.ksplits, &a.steps_total, &a.steps_per_split);
  a.OH = p->out_h; a.OW = p->out_w; a.wblocks = (p->out_w + kWgPix - 1) / kWgPix;
  a.sh = p->stride_h; a.sw = p->stride_w; a.nb = (2 * p->cin) / 64; a.N = 2 * p->cin;
  a.f16 = p->dtype == DCS_F16 ? 1 : 0;
  memcpy(a.dy, p->dy_off, sizeof(a.dy));
  memcpy(a.dx, p->dx_off, sizeof(a.dx));
  a.partial = reinterpret_cast<float*>(p->workspace);
  CUtensorMap tmY, tmX;
  if (int e = wg_map(&tmY, p->dy, a.f16, 2 * p->cout, p->out_w, p->out_h, p->batch, 1)) return e;
  if (int e = wg_map(&tmX, p->x, a.f16, 2 * p->cin, p->in_w, p->in_h, p->batch, p->stride_w)) return e;
  const size_t smem = 1024 + (size_t)kWgStages * (2 + a.nb) * kWgBoxBytes + sizeof(WgBars);
  DCS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t s = (cudaStream_t)stream;
  wgrad_tc_kernel<<<a.ksplits * a.n_taps * a.m_tiles, kWgThreads, smem, s>>>(tmY, tmX, a);
  DCS_LAUNCHED();
  const int n_elem = p->cout * p->cin * p->ntaps;
  wgrad_fold_kernel<<<std::min((n_elem + 255) / 256, 4 * num_sms()), 256, 0, s>>>(a.partial, a.ksplits, a.n_taps, p->cout, p->cin, p->dw_r, p->dw_i);
  DCS_LAUNCHED();
  return 0;
}

// ---- the same GEMM behind dcs_wgrad's interface (real channel counts, channel-slice pitches, any k2 <= 256 / n2: operand boxes beyond
//      the real channels are zero-filled by TMA, so M = n2 is padded to 128-row tiles and N = k2 to 64-column boxes)
static void wg16_shape(const dcs_wgrad16_params* p, dcs_cwgrad_params* q) {
  memset(q, 0, sizeof(*q));
  q->batch = p->batch; q->in_h = p->in_h; q->in_w = p->in_w; q->out_h = p->out_h; q->out_w = p->out_w;
  q->cin = (p->k2 + 1) / 2; q->cout = (p->n2 + 1) / 2; q->stride_h = p->stride_h; q->stride_w = p->stride_w; q->ntaps = p->ntaps;
}
extern "C" int64_t dcs_wgrad_tc16_workspace_bytes(const dcs_wgrad16_params* p) {
  if (!p || p->batch <= 0 || p->k2 <= 0 || p->n2 <= 0 || p->ntaps <= 0) return -1;
  dcs_cwgrad_params q;
  wg16_shape(p, &q);
  int ks, st, per;
  wg_split(&q, &ks, &st, &per);
  const int Mp = (p->n2 + 127) / 128 * 128, Np = (p->k2 + 63) / 64 * 64;
  return (int64_t)ks * p->ntaps * Mp * Np * (int64_t)sizeof(float);
}
extern "C" int dcs_wgrad_tc16(const dcs_wgrad16_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->dy && p->dwp && p->workspace, "dcs_wgrad_tc16: null pointer");
  DCS_REQUIRE(is_h16(p->dtype), "dcs_wgrad_tc16: x / dy must be DCS_F16 or DCS_BF16 storage");
  DCS_REQUIRE(p->batch > 0 && p->in_h > 0 && p->in_w > 0 && p->out_h > 0 && p->out_w > 0 && p->k2 > 0 && p->k2 <= 256 && p->n2 > 0, "dcs_wgrad_tc16: bad shape (k2 <= 256)");
  DCS_REQUIRE(p->x_pitch >= p->k2 && p->dy_pitch >= p->n2 && p->x_pitch % 8 == 0 && p->dy_pitch % 8 == 0, "dcs_wgrad_tc16: pitches must be multiples of 8 elements");
  DCS_REQUIRE(p->ntaps >= 1 && p->ntaps <= DCS_MAX_TAPS && (p->stride_w == 1 || p->stride_w == 2) && p->stride_h >= 1, "dcs_wgrad_tc16: bad taps / stride");
  DCS_REQUIRE(((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->dy % 16 == 0), "dcs_wgrad_tc16: pointers must be 16-byte aligned");
  DCS_REQUIRE(p->workspace_bytes >= dcs_wgrad_tc16_workspace_bytes(p), "dcs_wgrad_tc16: workspace too small");
  dcs_cwgrad_params q;
  wg16_shape(p, &q);
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.n_taps = p->ntaps; a.m_tiles = (p->n2 + 127) / 128;
  wg_split(&q, &a.ksplits, &a.steps_total, &a.steps_per_split);
  a.OH = p->out_h; a.OW = p->out_w; a.wblocks = (p->out_w + kWgPix - 1) / kWgPix;
  a.sh = p->stride_h; a.sw = p->stride_w; a.nb = (p->k2 + 63) / 64; a.N = a.nb * 64;
  a.f16 = p->dtype == DCS_F16 ? 1 : 0;
  memcpy(a.dy, p->dy_off, sizeof(a.dy));
  memcpy(a.dx, p->dx_off, sizeof(a.dx));
  a.partial = reinterpret_cast<float*>(p->workspace);
  CUtensorMap tmY, tmX;
  if (int e = wg_map(&tmY, p->dy, a.f16, p->n2, p->out_w, p->out_h, p->batch, 1, p->dy_pitch)) return e;
  if (int e = wg_map(&tmX, p->x, a.f16, p->k2, p->in_w, p->in_h, p->batch, p->stride_w, p->x_pitch)) return e;
  const size_t smem = 1024 + (size_t)kWgStages * (2 + a.nb) * kWgBoxBytes + sizeof(WgBars);
  DCS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t s = (cudaStream_t)stream;
  wgrad_tc_kernel<<<a.ksplits * a.n_taps * a.m_tiles, kWgThreads, smem, s>>>(tmY, tmX, a);
  DCS_LAUNCHED();
  const int n_elem = p->ntaps * p->k2 * p->n2;
  wgrad_reduce_generic_kernel<<<std::min((n_elem + 255) / 256, 8 * num_sms()), 256, 0, s>>>(a.partial, a.ksplits, a.n_taps, a.m_tiles * 128, a.N, p->k2, p->n2,
                                                                                            p->dwp, p->dwp_tap_stride > 0 ? p->dwp_tap_stride : (int64_t)p->k2 * p->n2);
  DCS_LAUNCHED();
  return 0;
}
