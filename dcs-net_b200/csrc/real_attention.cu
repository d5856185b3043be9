// real_attention.cu — RealChannelAttention + RealSpatialAttention of the real network path (SURVEY 8f rank 1;
// /root/reference/r_network.py:8-40, applied at 152-156 and 163-165):
//     gate_c = sigmoid(W2 relu(W1 maxpool_hw(x)))          (only the max-pool branch reaches the output, line 24)
//     u      = gate_c * x
//     gate_s = sigmoid(conv7x7([mean_c u, max_c u]))        (zero padding 3, no bias)
//     y      = gate_s * u
// on channels-last REAL tensors (B, H, W, C) — the same memory as the channel-pair tensors (B, H, W, C/2, 2) of the conv
// kernels.  First, straightforward fp32 / bf16 version (four small kernels, x read three times): correctness first, the
// streaming single-pass form of attention_stream.cu is the follow-up.
#include "common.cuh"

namespace dcs {

template <typename T> __device__ __forceinline__ float ldr(const T* p, int64_t i);
template <> __device__ __forceinline__ float ldr<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldr<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
template <> __device__ __forceinline__ float ldr<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <typename T> __device__ __forceinline__ void str(T* p, int64_t i, float v);
template <> __device__ __forceinline__ void str<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void str<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void str<__half>(__half* p, int64_t i, float v) { p[i] = from_float<__half>(v); }

// 1. per-(image, chunk, channel) maximum over a chunk of pixels -> scratch[b][chunk][c]
template <typename T>
__global__ void __launch_bounds__(256) real_chan_max_kernel(const T* __restrict__ x, float* __restrict__ scratch, int hw, int C, int ppc) {
  __shared__ float red[256];
  const int b = blockIdx.y, chunk = blockIdx.x, n_chunks = gridDim.x;
  const int lanes = 256 / C, c = threadIdx.x % C, pl = threadIdx.x / C;
  const int p0 = chunk * ppc, p1 = min(p0 + ppc, hw);
  float m = -INFINITY;
  const T* xb = x + (int64_t)b * hw * C;
  for (int p = p0 + pl; p < p1; p += lanes) m = fmaxf(m, ldr<T>(xb, (int64_t)p * C + c));
  red[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x < C) {
    for (int l = 1; l < lanes; ++l) m = fmaxf(m, red[l * C + threadIdx.x]);
    scratch[((int64_t)b * n_chunks + chunk) * C + threadIdx.x] = m;
  }
}

// 2. gate[b][c] = sigmoid(W2 relu(W1 max)), W1 (R, C), W2 (C, R)
__global__ void __launch_bounds__(256) real_gate_kernel(const float* __restrict__ scratch, const float* __restrict__ w1,
                                                        const float* __restrict__ w2, float* __restrict__ gate, int n_chunks, int C, int R) {
  __shared__ float mx[256];
  __shared__ float hid[16];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < C) {
    float m = -INFINITY;
    for (int k = 0; k < n_chunks; ++k) m = fmaxf(m, scratch[((int64_t)b * n_chunks + k) * C + tid]);
    mx[tid] = m;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < R; r += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w1[r * C + c] * mx[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hid[r] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (tid < C) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += w2[tid * R + r] * hid[r];
    gate[(int64_t)b * C + tid] = sigmoidf_(s);
  }
}

// 3. stats[b][p] = (mean_c u, max_c u), u = gate_c * x; G = min(32, C) lanes per pixel
template <typename T>
__global__ void __launch_bounds__(256) real_spat_stats_kernel(const T* __restrict__ x, const float* __restrict__ gate,
                                                              float2* __restrict__ stats, int hw, int C, int G) {
  __shared__ float gs[256];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += 256) gs[c] = gate[(int64_t)b * C + c];
  __syncthreads();
  const int sub = threadIdx.x % G, grp = threadIdx.x / G, groups = 256 / G;
  const T* xb = x + (int64_t)b * hw * C;
  const float invC = 1.f / (float)C;
  for (int pbase = blockIdx.x * groups; pbase < hw; pbase += gridDim.x * groups) {   // CTA-uniform trip count (shuffles)
    const int p = pbase + grp;
    float s = 0.f, m = -INFINITY;
    if (p < hw)
      for (int c = sub; c < C; c += G) {
        const float u = gs[c] * ldr<T>(xb, (int64_t)p * C + c);
        s += u; m = fmaxf(m, u);
      }
    for (int o = G >> 1; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if (sub == 0 && p < hw) stats[(int64_t)b * hw + p] = make_float2(s * invC, m);
  }
}

// 4. one warp per pixel: 7x7 conv over the two statistics planes (zero padding), sigmoid, y = gate_s * gate_c * x
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) real_spat_apply_kernel(const TI* __restrict__ x, const float* __restrict__ gate,
                                                              const float2* __restrict__ stats, const float* __restrict__ w7,
                                                              TO* __restrict__ y, int H, int W, int C) {
  __shared__ float gs[256];
  __shared__ float wq[98];      // [mean | max] x 49 taps (conv1.weight (1, 2, 7, 7))
  const int b = blockIdx.y, hw = H * W;
  for (int c = threadIdx.x; c < C; c += 256) gs[c] = gate[(int64_t)b * C + c];
  for (int i = threadIdx.x; i < 98; i += 256) wq[i] = w7[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float2* sb = stats + (int64_t)b * hw;
  for (int p = blockIdx.x * 8 + warp; p < hw; p += gridDim.x * 8) {
    const int py = p / W, px = p - py * W;
    float acc = 0.f;
    for (int i = lane; i < 49; i += 32) {
      const int yy = py + i / 7 - 3, xx = px + i % 7 - 3;
      if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
        const float2 st = sb[yy * W + xx];
        acc += wq[i] * st.x + wq[49 + i] * st.y;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float s = sigmoidf_(acc);
    const int64_t base = ((int64_t)b * hw + p) * C;
    for (int c = lane; c < C; c += 32) str<TO>(y, base + c, s * (gs[c] * ldr<TI>(x, base + c)));
  }
}

}  // namespace dcs

using namespace dcs;

extern "C" int64_t dcs_real_attention_workspace_bytes(int batch, int h, int w, int channels) {
  if (batch <= 0 || h <= 0 || w <= 0 || channels <= 0) return -1;
  const int64_t hw = (int64_t)h * w;
  return ((int64_t)batch * 64 * channels + (int64_t)batch * channels) * 4 + batch * hw * 8;   // partial maxima, gate, statistics
}

extern "C" int dcs_real_attention_fwd(const dcs_real_attention_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->w1 && p->w2 && p->w7 && p->workspace, "dcs_real_attention_fwd: null pointer");
  const int C = p->channels, R = p->reduced;
  DCS_REQUIRE(p->batch > 0 && p->batch <= 65535 && p->h > 0 && p->w > 0, "dcs_real_attention_fwd: bad shape");
  DCS_REQUIRE(C >= 1 && C <= 256 && (C & (C - 1)) == 0 && R >= 1 && R <= 16, "dcs_real_attention_fwd: channels must be a power of two <= 256, reduced <= 16");
  DCS_REQUIRE(p->workspace_bytes >= dcs_real_attention_workspace_bytes(p->batch, p->h, p->w, C), "dcs_real_attention_fwd: workspace too small");
  const int hw = p->h * p->w;
  float* scratch = reinterpret_cast<float*>(p->workspace);
  float* gate = scratch + (int64_t)p->batch * 64 * C;
  float2* stats = reinterpret_cast<float2*>(gate + (int64_t)p->batch * C);
  cudaStream_t s = (cudaStream_t)stream;
  const int lanes = 256 / C;
  int n_chunks = std::min(64, std::max(1, hw / (lanes * 8)));
  const int ppc = (hw + n_chunks - 1) / n_chunks;
  n_chunks = (hw + ppc - 1) / ppc;
  DCS_REQUIRE(is_dtype(p->dtype), "dcs_real_attention_fwd: bad dtype");
  const bool bf = p->dtype == DCS_BF16, hf = p->dtype == DCS_F16;
  if (bf) real_chan_max_kernel<__nv_bfloat16><<<dim3(n_chunks, p->batch), 256, 0, s>>>((const __nv_bfloat16*)p->x, scratch, hw, C, ppc);
  else if (hf) real_chan_max_kernel<__half><<<dim3(n_chunks, p->batch), 256, 0, s>>>((const __half*)p->x, scratch, hw, C, ppc);
  else real_chan_max_kernel<float><<<dim3(n_chunks, p->batch), 256, 0, s>>>((const float*)p->x, scratch, hw, C, ppc);
  DCS_LAUNCHED();
  real_gate_kernel<<<p->batch, 256, 0, s>>>(scratch, p->w1, p->w2, gate, n_chunks, C, R);
  DCS_LAUNCHED();
  const int G = std::min(32, C), groups = 256 / G;
  const int ctas = std::max(1, std::min((hw + groups - 1) / groups, 16 * num_sms() / p->batch + 1));
  if (bf) real_spat_stats_kernel<__nv_bfloat16><<<dim3(ctas, p->batch), 256, 0, s>>>((const __nv_bfloat16*)p->x, gate, stats, hw, C, G);
  else if (hf) real_spat_stats_kernel<__half><<<dim3(ctas, p->batch), 256, 0, s>>>((const __half*)p->x, gate, stats, hw, C, G);
  else real_spat_stats_kernel<float><<<dim3(ctas, p->batch), 256, 0, s>>>((const float*)p->x, gate, stats, hw, C, G);
  DCS_LAUNCHED();
  const int actas = std::max(1, std::min((hw + 7) / 8, 32 * num_sms() / p->batch + 1));
  if (bf) real_spat_apply_kernel<__nv_bfloat16, __nv_bfloat16><<<dim3(actas, p->batch), 256, 0, s>>>((const __nv_bfloat16*)p->x, gate, stats, p->w7, (__nv_bfloat16*)p->y, p->h, p->w, C);
  else if (hf) real_spat_apply_kernel<__half, __half><<<dim3(actas, p->batch), 256, 0, s>>>((const __half*)p->x, gate, stats, p->w7, (__half*)p->y, p->h, p->w, C);
  else real_spat_apply_kernel<float, float><<<dim3(actas, p->batch), 256, 0, s>>>((const float*)p->x, gate, stats, p->w7, (float*)p->y, p->h, p->w, C);
  DCS_LAUNCHED();
  return 0;
}
