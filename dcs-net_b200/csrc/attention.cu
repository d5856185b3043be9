// attention.cu — CBAM-style complex channel / spatial attention (bandwidth-bound passes).
//
// Reference: ComplexChannelAttention /root/reference/c_network.py:53-69 (pools network_functions.py:114-138 —
// the "max" pool is an average pool, so gate = sigmoid_c(2*fc(avg))), ComplexSpatialAttention c_network.py:71-84,
// ComplexSigmoid network_functions.py:107-112; applied with full complex products at c_network.py:208-211,219-220.
//
// Passes over a (B,H,W,C) channels-last complex tensor x:
//   1. chan_pool : sums[b][c]  = sum_hw x                      (read x once)          [fused into the conv epilogue
//   2. chan_gate : g_c[b][c]   = sigmoid_c(2 W2 crelu(W1 avg)) (tiny)                  on the tcgen05 path]
//   3. spat_stats: st[b][h][w] = {mean_c(g_c x), max_c Re, max_c Im}   (read x once, write 16 B / pixel)
//   4. spat_apply: y = sigmoid_c(conv7x7(st)) * (g_c x)                (read x once, write y once)
#include <string.h>
#include <algorithm>
#include "common.cuh"

namespace dcs {

// 16-byte vector access: VEC complex elements per lane (4 for bf16, 2 for fp32)
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void ld(const float* p, int64_t cidx, float2 (&v)[2]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p + 2 * cidx));
    v[0] = make_float2(q.x, q.y); v[1] = make_float2(q.z, q.w);
  }
  static __device__ __forceinline__ void st(float* p, int64_t cidx, const float2 (&v)[2]) {
    *reinterpret_cast<float4*>(p + 2 * cidx) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  }
};
template <typename T> struct Vec16H {   // 16-bit storage (fp16 / bf16): 4 complex per 16 bytes
  static constexpr int N = 4;
  static __device__ __forceinline__ void ld(const T* p, int64_t cidx, float2 (&v)[4]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + 2 * cidx));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = unpack_h2<T>(w[i]);
  }
  static __device__ __forceinline__ void st(T* p, int64_t cidx, const float2 (&v)[4]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = pack_h2<T>(v[i].x, v[i].y);
    *reinterpret_cast<uint4*>(p + 2 * cidx) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec16<__nv_bfloat16> : Vec16H<__nv_bfloat16> {};
template <> struct Vec16<__half> : Vec16H<__half> {};

// ---------------------------------------------------------------- 1. global average pool (sums)
template <typename T>
__global__ void __launch_bounds__(256) chan_pool_kernel(const T* __restrict__ x, long long* __restrict__ sums, int hw, int C,
                                                        int pix_per_cta) {
  __shared__ float2 red[256];
  const int b = blockIdx.y;
  const int lanes = 256 / C;            // pixel lanes (C <= 256, power of two)
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, hw);
  float2 acc = make_float2(0.f, 0.f);
  if (pl < lanes) {
    const T* xb = x + (int64_t)b * hw * C * 2;
    for (int p = p0 + pl; p < p1; p += lanes) {
      const float2 v = Elem<T>::ldc(xb, (int64_t)p * C + c);
      acc.x += v.x; acc.y += v.y;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float2 s = make_float2(0.f, 0.f);
    for (int l = 0; l < lanes; ++l) { const float2 v = red[l * C + threadIdx.x]; s.x += v.x; s.y += v.y; }
    pool_add(sums + ((int64_t)b * C + threadIdx.x) * 2 + 0, s.x);
    pool_add(sums + ((int64_t)b * C + threadIdx.x) * 2 + 1, s.y);
  }
}

// per-(image, pair, slot) maxima in the DCS_POOL_MAX encoding (the real path's AdaptiveMaxPool2d(1))
template <typename T>
__global__ void __launch_bounds__(256) chan_max_kernel(const T* __restrict__ x, long long* __restrict__ maxima, int hw, int C,
                                                       int pix_per_cta) {
  __shared__ float2 red[256];
  const int b = blockIdx.y;
  const int lanes = 256 / C;
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, hw);
  float2 acc = make_float2(-INFINITY, -INFINITY);
  if (pl < lanes) {
    const T* xb = x + (int64_t)b * hw * C * 2;
    for (int p = p0 + pl; p < p1; p += lanes) {
      const float2 v = Elem<T>::ldc(xb, (int64_t)p * C + c);
      acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float2 s = make_float2(-INFINITY, -INFINITY);
    for (int l = 0; l < lanes; ++l) { const float2 v = red[l * C + threadIdx.x]; s.x = fmaxf(s.x, v.x); s.y = fmaxf(s.y, v.y); }
    pool_max(maxima + ((int64_t)b * C + threadIdx.x) * 2 + 0, s.x);
    pool_max(maxima + ((int64_t)b * C + threadIdx.x) * 2 + 1, s.y);
  }
}

// ---------------------------------------------------------------- 2. the squeeze/excite MLP on (B, C) complex
__global__ void __launch_bounds__(128) chan_gate_kernel(const dcs_chan_gate_params p) {
  __shared__ float2 avg[256];
  __shared__ float2 hid[16];
  const int b = blockIdx.x, C = p.channels, R = p.reduced;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long long* sm = reinterpret_cast<const long long*>(p.sums);
    avg[c] = make_float2(pool_mean(sm, ((int64_t)b * C + c) * 2, p.inv_hw), pool_mean(sm, ((int64_t)b * C + c) * 2 + 1, p.inv_hw));
  }
  __syncthreads();
  // hidden r: one warp per r (C <= 256), complex 1x1 conv without bias then ComplexReLU
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += blockDim.x >> 5) {
    float re = 0.f, im = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float wr = p.w1_r[r * C + c], wi = p.w1_i[r * C + c];
      re += wr * avg[c].x - wi * avg[c].y;
      im += wr * avg[c].y + wi * avg[c].x;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float re = 0.f, im = 0.f;
    for (int r = 0; r < R; ++r) {
      const float wr = p.w2_r[c * R + r], wi = p.w2_i[c * R + r];
      re += wr * hid[r].x - wi * hid[r].y;
      im += wr * hid[r].y + wi * hid[r].x;
    }
    // fc(avg) + fc("max" = avg) = 2 fc(avg)
    p.gate[((int64_t)b * C + c) * 2 + 0] = sigmoidf_(2.f * re);
    p.gate[((int64_t)b * C + c) * 2 + 1] = sigmoidf_(2.f * im);
  }
}

// ---------------------------------------------------------------- 3. per-pixel channel statistics of u = g_c * x
struct GateMlp {   // optional fused ComplexChannelAttention MLP (dcs_chan_gate) in the statistics kernel
  const long long* sums; float inv_hw; int reduced;
  const float *w1_r, *w1_i, *w2_r, *w2_i;
  float* gate_out;
};

template <typename T>
__global__ void __launch_bounds__(256) spat_stats_kernel(const T* __restrict__ x, const float* __restrict__ gate,
                                                         float4* __restrict__ stats, int hw, int C, int G, const GateMlp mlp) {
  // G lanes cooperate on one pixel; each lane strides over the channels in 16-byte vectors (G = min(32, C / VEC))
  constexpr int V = Vec16<T>::N;
  __shared__ float2 gs[256];
  __shared__ float2 avg[256];
  __shared__ float2 hid[16];
  const int b = blockIdx.y;
  if (mlp.sums) {
    // gate = sigmoid_c(2 W2 crelu(W1 avg)) recomputed by every CTA of the image (C*R complex MACs); CTA 0 publishes it
    const int R = mlp.reduced;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      avg[c] = make_float2(pool_mean(mlp.sums, ((int64_t)b * C + c) * 2, mlp.inv_hw), pool_mean(mlp.sums, ((int64_t)b * C + c) * 2 + 1, mlp.inv_hw));
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < R; r += 8) {
      float re = 0.f, im = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float wr = mlp.w1_r[r * C + c], wi = mlp.w1_i[r * C + c];
        re += wr * avg[c].x - wi * avg[c].y;
        im += wr * avg[c].y + wi * avg[c].x;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
      if (lane == 0) hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float re = 0.f, im = 0.f;
      for (int r = 0; r < R; ++r) {
        const float wr = mlp.w2_r[c * R + r], wi = mlp.w2_i[c * R + r];
        re += wr * hid[r].x - wi * hid[r].y;
        im += wr * hid[r].y + wi * hid[r].x;
      }
      const float2 gv = make_float2(sigmoidf_(2.f * re), sigmoidf_(2.f * im));
      gs[c] = gv;
      if (blockIdx.x == 0 && mlp.gate_out) reinterpret_cast<float2*>(mlp.gate_out)[(int64_t)b * C + c] = gv;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      gs[c] = gate ? reinterpret_cast<const float2*>(gate)[(int64_t)b * C + c] : make_float2(1.f, 0.f);
  }
  __syncthreads();
  const int sub = threadIdx.x % G, grp = threadIdx.x / G, groups = 256 / G;
  const T* xb = x + (int64_t)b * hw * C * 2;
  const float invC = 1.f / (float)C;
  for (int pbase = blockIdx.x * groups; pbase < hw; pbase += gridDim.x * groups) {  // CTA-uniform trip count (shuffles)
    const int p = pbase + grp;
    float sr = 0.f, si = 0.f, mr = -INFINITY, mi = -INFINITY;
    if (p < hw) {
      for (int c = sub * V; c < C; c += G * V) {
        float2 v[V];
        Vec16<T>::ld(xb, (int64_t)p * C + c, v);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float2 u = cmul(gs[c + e], v[e]);
          sr += u.x; si += u.y;
          mr = fmaxf(mr, u.x); mi = fmaxf(mi, u.y);
        }
      }
    }
    for (int o = G >> 1; o; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o);
      mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    }
    if (sub == 0 && p < hw) stats[(int64_t)b * hw + p] = make_float4(sr * invC, si * invC, mr, mi);
  }
}

// ---------------------------------------------------------------- 4. 7x7 complex conv on the stats, sigmoid, apply
constexpr int kSaK = 7, kSaR = 3;  // tile = kSaTH x kSaTW pixels, kSaTH in {4, 16}, kSaTW in {16, 64} (template)

template <typename TI, typename TO, int kSaTH, int kSaTW>
__global__ void __launch_bounds__(256) spat_apply_kernel(const TI* __restrict__ x, const float* __restrict__ gate,
                                                         const float4* __restrict__ stats, const float* __restrict__ w7,
                                                         TO* __restrict__ y, float2* __restrict__ gate_out, int H, int W, int C) {
  __shared__ float4 st[kSaTH + 2 * kSaR][kSaTW + 2 * kSaR];
  __shared__ float2 sg[kSaTH][kSaTW];
  __shared__ float2 gs[256];
  __shared__ float4 wq[49];
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * kSaTH, x0 = blockIdx.x * kSaTW;
  // w7 = [conv_r (1,2,7,7) | conv_i (1,2,7,7)]: regroup per tap as (Wr[mean], Wr[max], Wi[mean], Wi[max])
  for (int i = threadIdx.x; i < 49; i += 256) wq[i] = make_float4(w7[i], w7[49 + i], w7[98 + i], w7[147 + i]);
  for (int c = threadIdx.x; c < C; c += 256)
    gs[c] = gate ? reinterpret_cast<const float2*>(gate)[(int64_t)b * C + c] : make_float2(1.f, 0.f);
  for (int i = threadIdx.x; i < (kSaTH + 2 * kSaR) * (kSaTW + 2 * kSaR); i += 256) {
    const int r = i / (kSaTW + 2 * kSaR), cidx = i % (kSaTW + 2 * kSaR);
    const int yy = y0 + r - kSaR, xx = x0 + cidx - kSaR;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) v = stats[((int64_t)b * H + yy) * W + xx];
    st[r][cidx] = v;
  }
  __syncthreads();
  if (threadIdx.x < (kSaTH / 4) * kSaTW) {
    // ComplexConv2d(2,1,7,padding=3,bias=False) then ComplexSigmoid.  A thread owns 4 VERTICALLY adjacent pixels of one
    // tile column: consecutive lanes read consecutive 16-byte statistics (conflict-free LDS.128 — four horizontally
    // adjacent pixels per thread put the lanes 64 bytes apart: a 4-way bank conflict, 58 M conflicts per launch at
    // C = 8 in the r01b profile), each loaded value feeds up to 4 x 7 taps from registers, complex MACs as packed
    // fp32x2 FMAs.
    const int col = threadIdx.x % kSaTW, r = (threadIdx.x / kSaTW) * 4;
    float2 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
    for (int kx = 0; kx < kSaK; ++kx) {
      float4 sv[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) sv[j] = st[r + j][col + kx];
#pragma unroll
      for (int ky = 0; ky < kSaK; ++ky) {
        const float4 wv = wq[ky * 7 + kx];  // (Wr[mean], Wr[max], Wi[mean], Wi[max])
        const float2 w_mean = make_float2(wv.x, wv.z), w_mean_j = make_float2(-wv.z, wv.x);   // w and j*w
        const float2 w_max = make_float2(wv.y, wv.w), w_max_j = make_float2(-wv.w, wv.y);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 sq = sv[q + ky];     // (mean.re, mean.im, max.re, max.im)
          ffma2(acc[q], w_mean, make_float2(sq.x, sq.x));
          ffma2(acc[q], w_mean_j, make_float2(sq.y, sq.y));
          ffma2(acc[q], w_max, make_float2(sq.z, sq.z));
          ffma2(acc[q], w_max_j, make_float2(sq.w, sq.w));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 gv = make_float2(sigmoidf_(acc[q].x), sigmoidf_(acc[q].y));
      sg[r + q][col] = gv;
      if (gate_out && y0 + r + q < H && x0 + col < W) gate_out[((int64_t)b * H + y0 + r + q) * W + x0 + col] = gv;
    }
  }
  if (!y) return;
  __syncthreads();
  const int wvalid = min(kSaTW, W - x0), hvalid = min(kSaTH, H - y0);
  if constexpr (sizeof(TI) == sizeof(TO)) {  // same storage type: 16-byte vectors, one flat loop over the tile
    constexpr int V = Vec16<TI>::N;
    const int clog2 = 31 - __clz(C);         // C is a power of two (checked on the host)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warps stream tile rows (contiguous wvalid*C complex values); short tiles split each row over several warps
    auto apply_row = [&](int r, int iv0, int step) {
      const int64_t base = (((int64_t)b * H + y0 + r) * W + x0) * C;
#pragma unroll 4
      for (int iv = iv0; iv < wvalid * C; iv += step) {
        const int px = iv >> clog2, c = iv & (C - 1);
        float2 v[V];
        Vec16<TI>::ld(x, base + iv, v);
        const float2 g = sg[r][px];
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = cmul(g, cmul(gs[c + e], v[e]));
        Vec16<TO>::st(y, base + iv, v);
      }
    };
    if (hvalid >= 8) {
      for (int r = warp; r < hvalid; r += 8) apply_row(r, lane * V, 32 * V);
    } else {
      const int segs = 8 / hvalid, seg = warp / hvalid;
      if (seg < segs) apply_row(warp - seg * hvalid, (seg * 32 + lane) * V, segs * 32 * V);
    }
  } else {
    for (int r = 0; r < hvalid; ++r) {
      const int64_t base = (((int64_t)b * H + y0 + r) * W + x0) * C;
      for (int i = threadIdx.x; i < wvalid * C; i += 256) {
        const int px = i / C, c = i - px * C;
        const float2 u = cmul(gs[c], Elem<TI>::ldc(x, base + i));
        Elem<TO>::stc(y, base + i, cmul(sg[r][px], u));
      }
    }
  }
}

// ---------------------------------------------------------------- fused: channel gate MLP + spatial attention, ONE pass
// y = SA(u) * u with u = CA(x) * x  (c_network.py:208-211 / 219-220) for one (TH x TW) pixel tile per CTA:
//   0. the squeeze/excite MLP on this image's pooled sums (recomputed per CTA: C*R complex MACs, negligible)
//   1. x tile + 3-pixel halo -> shared memory (the only HBM read of x; halo re-reads are L2 hits)
//   2. per-pixel channel statistics of u over tile + halo -> shared (zero outside the image = the conv's zero padding)
//   3. 7x7 complex conv + ComplexSigmoid -> spatial gate of the inner tile
//   4. y = gate_s * gate_c * x from the shared tile -> HBM
// HBM traffic: x once in, y once out (the separate stats / apply kernels read x twice and round-trip 16 B / pixel of
// statistics).  TH, TW are runtime (TW % 4 == 0, TW * TH <= 1024); channels C = 1 << clog2.
constexpr int kFaR = 3, kFaThreads = 256;

struct FusedAttArgs {
  const void* x; void* y;
  const long long* sums; float inv_hw;
  const float *w1_r, *w1_i, *w2_r, *w2_i, *w7;
  int H, W, C, clog2, R, TH, TW;
};

template <typename TI, typename TO>
__global__ void __launch_bounds__(kFaThreads, 2) attention_fused_kernel(const FusedAttArgs a) {
  extern __shared__ __align__(16) unsigned char fa_smem[];
  constexpr int VI = Vec16<TI>::N;                 // complex elements per 16-byte vector of the input type
  const int C = a.C, TH = a.TH, TW = a.TW;
  const int PH = TH + 2 * kFaR, PW = TW + 2 * kFaR;
  TI* xs = reinterpret_cast<TI*>(fa_smem);                                             // [PH][PW][C][2]
  float4* st = reinterpret_cast<float4*>(fa_smem + (size_t)PH * PW * C * 2 * sizeof(TI));   // [PH][PW]
  float2* sg = reinterpret_cast<float2*>(st + PH * PW);                                  // [TH][TW]
  float2* gs = sg + TH * TW;                                                             // [C] channel gate
  float2* avg = gs + C;                                                                  // [C]
  float2* hid = avg + C;                                                                 // [16]
  float4* wq = reinterpret_cast<float4*>(hid + 16);                                      // [49]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const TI* xb = reinterpret_cast<const TI*>(a.x) + (int64_t)b * a.H * a.W * C * 2;

  // ---- 1. issue the tile loads first (they overlap the gate MLP): 16-byte vectors, zero outside the image
  const int vec_per_px = C / VI;
  const int n_vec = PH * PW * vec_per_px;
  const int vlog2 = a.clog2 - (VI == 4 ? 2 : 1);
  for (int i = tid; i < n_vec; i += kFaThreads) {
    const int px = i >> vlog2, cv = i & (vec_per_px - 1);
    const int r = px / PW, c = px - r * PW;
    const int yy = y0 + r - kFaR, xx = x0 + c - kFaR;
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W)
      v = __ldg(reinterpret_cast<const uint4*>(xb + (((int64_t)yy * a.W + xx) * C + cv * VI) * 2));
    reinterpret_cast<uint4*>(xs)[i] = v;
  }
  // ---- 0. channel gate (ComplexChannelAttention: sigmoid_c(2 W2 crelu(W1 avg)))
  for (int i = tid; i < 49; i += kFaThreads) wq[i] = make_float4(a.w7[i], a.w7[49 + i], a.w7[98 + i], a.w7[147 + i]);
  for (int c = tid; c < C; c += kFaThreads)
    avg[c] = make_float2(pool_mean(a.sums, ((int64_t)b * C + c) * 2, a.inv_hw), pool_mean(a.sums, ((int64_t)b * C + c) * 2 + 1, a.inv_hw));
  __syncthreads();
  for (int r = warp; r < a.R; r += kFaThreads / 32) {
    float re = 0.f, im = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float wr = a.w1_r[r * C + c], wi = a.w1_i[r * C + c];
      re += wr * avg[c].x - wi * avg[c].y;
      im += wr * avg[c].y + wi * avg[c].x;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
  }
  __syncthreads();
  for (int c = tid; c < C; c += kFaThreads) {
    float re = 0.f, im = 0.f;
    for (int r = 0; r < a.R; ++r) {
      const float wr = a.w2_r[c * a.R + r], wi = a.w2_i[c * a.R + r];
      re += wr * hid[r].x - wi * hid[r].y;
      im += wr * hid[r].y + wi * hid[r].x;
    }
    gs[c] = make_float2(sigmoidf_(2.f * re), sigmoidf_(2.f * im));
  }
  __syncthreads();
  // ---- 2. channel statistics of u = g_c * x per pixel of tile + halo; G lanes share a pixel
  {
    const int G = min(32, vec_per_px), glog2 = 31 - __clz(G);
    const int sub = tid & (G - 1), grp = tid >> glog2, groups = kFaThreads >> glog2;
    const float invC = 1.f / (float)C;
    for (int pbase = 0; pbase < PH * PW; pbase += groups) {   // CTA-uniform trip count (shuffles)
      const int p = pbase + grp;
      float sr = 0.f, si = 0.f, mr = -INFINITY, mi = -INFINITY;
      if (p < PH * PW) {
        for (int c = sub * VI; c < C; c += G * VI) {
          float2 v[VI];
          const uint4 q = *reinterpret_cast<const uint4*>(xs + ((int64_t)p * C + c) * 2);
          if constexpr (VI == 4) {
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = unpack_h2<TI>(w4[e]);
          } else {
            v[0] = make_float2(__uint_as_float(q.x), __uint_as_float(q.y));
            v[1] = make_float2(__uint_as_float(q.z), __uint_as_float(q.w));
          }
#pragma unroll
          for (int e = 0; e < VI; ++e) {
            const float2 u = cmul(gs[c + e], v[e]);
            sr += u.x; si += u.y;
            mr = fmaxf(mr, u.x); mi = fmaxf(mi, u.y);
          }
        }
      }
      for (int o = G >> 1; o; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o);
        mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
      }
      if (sub == 0 && p < PH * PW) {
        const int r = p / PW, c = p - r * PW;
        const int yy = y0 + r - kFaR, xx = x0 + c - kFaR;
        const bool in = (unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W;   // zero padding of the 7x7 conv
        st[p] = in ? make_float4(sr * invC, si * invC, mr, mi) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __syncthreads();
  // ---- 3. ComplexConv2d(2,1,7,padding=3,bias=False) + ComplexSigmoid.  A thread owns NR vertically adjacent pixels of
  //         one tile column: consecutive lanes read consecutive 16-byte statistics (conflict-free LDS.128; owning
  //         horizontally adjacent pixels puts the lanes 64 bytes apart, a 4-way bank conflict), and each loaded
  //         value is reused by up to NR x 7 taps from registers.
  {
    const int NR = min(4, TH), twlog = 31 - __clz(TW);
    for (int t = tid; t < (TH / NR) * TW; t += kFaThreads) {
      const int col = t & (TW - 1), r = (t >> twlog) * NR;
      float2 acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        float4 sv[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) sv[j] = (j < NR + 6) ? st[(r + j) * PW + col + kx] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const float4 wv = wq[ky * 7 + kx];  // (Wr[mean], Wr[max], Wi[mean], Wi[max])
          const float2 w_mean = make_float2(wv.x, wv.z), w_mean_j = make_float2(-wv.z, wv.x);   // w, j*w
          const float2 w_max = make_float2(wv.y, wv.w), w_max_j = make_float2(-wv.w, wv.y);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 sq = sv[q + ky];     // (mean.re, mean.im, max.re, max.im)
            ffma2(acc[q], w_mean, make_float2(sq.x, sq.x));
            ffma2(acc[q], w_mean_j, make_float2(sq.y, sq.y));
            ffma2(acc[q], w_max, make_float2(sq.z, sq.z));
            ffma2(acc[q], w_max_j, make_float2(sq.w, sq.w));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < NR) sg[(r + q) * TW + col] = make_float2(sigmoidf_(acc[q].x), sigmoidf_(acc[q].y));
    }
  }
  __syncthreads();
  // ---- 4. y = gate_s * (gate_c * x), inner tile, from the shared copy
  {
    TO* yb = reinterpret_cast<TO*>(a.y) + (int64_t)b * a.H * a.W * C * 2;
    const int n_inner = TH * TW * vec_per_px;
    const int twlog = 31 - __clz(TW);
    for (int i = tid; i < n_inner; i += kFaThreads) {
      const int px = i >> vlog2, cv = i & (vec_per_px - 1);
      const int r = px >> twlog, c = px & (TW - 1);
      const int yy = y0 + r, xx = x0 + c;
      if (yy >= a.H || xx >= a.W) continue;
      const int p = (r + kFaR) * PW + c + kFaR;
      const uint4 q = *reinterpret_cast<const uint4*>(xs + ((int64_t)p * C + cv * VI) * 2);
      float2 v[VI];
      if constexpr (VI == 4) {
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = unpack_h2<TI>(w4[e]);
      } else {
        v[0] = make_float2(__uint_as_float(q.x), __uint_as_float(q.y));
        v[1] = make_float2(__uint_as_float(q.z), __uint_as_float(q.w));
      }
      const float2 g = sg[r * TW + c];
#pragma unroll
      for (int e = 0; e < VI; ++e) v[e] = cmul(g, cmul(gs[cv * VI + e], v[e]));
      const int64_t o = ((int64_t)yy * a.W + xx) * C + cv * VI;
      if constexpr (sizeof(TO) == sizeof(TI)) {
        Vec16<TO>::st(yb, o, v);
      } else if constexpr (sizeof(TO) == 2) {          // fp32 in, 16-bit out: 2 complex = 8 bytes
        *reinterpret_cast<uint2*>(yb + 2 * o) = make_uint2(pack_h2<TO>(v[0].x, v[0].y), pack_h2<TO>(v[1].x, v[1].y));
      } else {                                         // bf16 in, fp32 out: 4 complex = 32 bytes
        *reinterpret_cast<float4*>(yb + 2 * o) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        *reinterpret_cast<float4*>(yb + 2 * o + 4) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
      }
    }
  }
}

}  // namespace dcs

using namespace dcs;

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" int dcs_chan_pool(const dcs_chan_pool_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->sums, "dcs_chan_pool: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->hw > 0 && pow2(p->channels) && p->channels <= 256, "dcs_chan_pool: channels must be a power of two <= 256");
  const int lanes = 256 / p->channels;
  int ctas = (p->hw + lanes * 8 - 1) / (lanes * 8);            // >= 8 pixels per lane
  const int cap = max(1, 8 * num_sms() / p->batch);
  ctas = min(ctas, cap);
  const int ppc = (p->hw + ctas - 1) / ctas;
  dim3 grid((p->hw + ppc - 1) / ppc, p->batch);
  cudaStream_t s = (cudaStream_t)stream;
  DCS_REQUIRE(is_dtype(p->dtype), "dcs_chan_pool: bad dtype");
  long long* sums = reinterpret_cast<long long*>(p->sums);
  if (p->dtype == DCS_BF16) chan_pool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)p->x, sums, p->hw, p->channels, ppc);
  else if (p->dtype == DCS_F16) chan_pool_kernel<__half><<<grid, 256, 0, s>>>((const __half*)p->x, sums, p->hw, p->channels, ppc);
  else chan_pool_kernel<float><<<grid, 256, 0, s>>>((const float*)p->x, sums, p->hw, p->channels, ppc);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_chan_max(const dcs_chan_pool_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->sums, "dcs_chan_max: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->hw > 0 && pow2(p->channels) && p->channels <= 256, "dcs_chan_max: channels must be a power of two <= 256");
  DCS_REQUIRE(is_dtype(p->dtype), "dcs_chan_max: bad dtype");
  const int lanes = 256 / p->channels;
  int ctas = (p->hw + lanes * 8 - 1) / (lanes * 8);
  ctas = min(ctas, max(1, 8 * num_sms() / p->batch));
  const int ppc = (p->hw + ctas - 1) / ctas;
  dim3 grid((p->hw + ppc - 1) / ppc, p->batch);
  cudaStream_t s = (cudaStream_t)stream;
  long long* mx = reinterpret_cast<long long*>(p->sums);
  if (p->dtype == DCS_BF16) chan_max_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)p->x, mx, p->hw, p->channels, ppc);
  else if (p->dtype == DCS_F16) chan_max_kernel<__half><<<grid, 256, 0, s>>>((const __half*)p->x, mx, p->hw, p->channels, ppc);
  else chan_max_kernel<float><<<grid, 256, 0, s>>>((const float*)p->x, mx, p->hw, p->channels, ppc);
  DCS_LAUNCHED();
  return 0;
}

namespace dcs {
__global__ void pool_mean_kernel(const long long* __restrict__ sums, float inv_hw, float* __restrict__ mean, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mean[i] = pool_mean(sums, i, inv_hw);
}
}  // namespace dcs

extern "C" int dcs_pool_mean(const int64_t* sums, float inv_hw, float* mean, int64_t n, void* stream) {
  DCS_REQUIRE(sums && mean && n > 0, "dcs_pool_mean: bad arguments");
  pool_mean_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(sums), inv_hw, mean, n);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_chan_gate(const dcs_chan_gate_params* p, void* stream) {
  DCS_REQUIRE(p && p->sums && p->gate && p->w1_r && p->w1_i && p->w2_r && p->w2_i, "dcs_chan_gate: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->channels > 0 && p->channels <= 256 && p->reduced > 0 && p->reduced <= 16, "dcs_chan_gate: bad shape");
  chan_gate_kernel<<<p->batch, 128, 0, (cudaStream_t)stream>>>(*p);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_spat_stats(const dcs_spat_stats_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->stats, "dcs_spat_stats: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->h > 0 && p->w > 0 && pow2(p->channels) && p->channels <= 256, "dcs_spat_stats: bad shape");
  const int hw = p->h * p->w;
  DCS_REQUIRE(is_dtype(p->dtype), "dcs_spat_stats: bad dtype");
  const int vec = is_h16(p->dtype) ? 4 : 2;
  DCS_REQUIRE(p->channels % vec == 0, "dcs_spat_stats: channels must be a multiple of %d", vec);
  const int G = min(32, p->channels / vec), groups = 256 / G;
  int ctas = (hw + groups - 1) / groups;
  ctas = min(ctas, max(1, 16 * num_sms() / p->batch));
  dim3 grid(ctas, p->batch);
  cudaStream_t s = (cudaStream_t)stream;
  GateMlp mlp;
  memset(&mlp, 0, sizeof(mlp));
  if (p->sums) {
    DCS_REQUIRE(p->w1_r && p->w1_i && p->w2_r && p->w2_i && p->reduced > 0 && p->reduced <= 16, "dcs_spat_stats: incomplete gate MLP operands");
    mlp.sums = reinterpret_cast<const long long*>(p->sums); mlp.inv_hw = 1.f / (float)hw; mlp.reduced = p->reduced;
    mlp.w1_r = p->w1_r; mlp.w1_i = p->w1_i; mlp.w2_r = p->w2_r; mlp.w2_i = p->w2_i; mlp.gate_out = p->gate_out;
  }
  if (p->dtype == DCS_BF16) spat_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)p->x, p->chan_gate, (float4*)p->stats, hw, p->channels, G, mlp);
  else if (p->dtype == DCS_F16) spat_stats_kernel<__half><<<grid, 256, 0, s>>>((const __half*)p->x, p->chan_gate, (float4*)p->stats, hw, p->channels, G, mlp);
  else spat_stats_kernel<float><<<grid, 256, 0, s>>>((const float*)p->x, p->chan_gate, (float4*)p->stats, hw, p->channels, G, mlp);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_spat_apply(const dcs_spat_apply_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->stats && p->w7 && (p->y || p->gate_out), "dcs_spat_apply: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->h > 0 && p->w > 0 && p->channels > 0 && p->channels <= 256, "dcs_spat_apply: bad shape");
  DCS_REQUIRE(p->batch <= 65535, "dcs_spat_apply: batch too large for grid.z");
  DCS_REQUIRE(pow2(p->channels) && p->channels >= 4, "dcs_spat_apply: channels must be a power of two >= 4");
  const int th = p->h >= 16 ? 16 : 4;
  // 64-pixel-wide tiles; wide-channel tensors have few pixels, so narrow 16-pixel tiles keep >= ~4 CTAs per SM busy
  const int64_t ctas64 = (int64_t)((p->w + 63) / 64) * ((p->h + th - 1) / th) * p->batch;
  const int tw = ctas64 < 4 * num_sms() ? 16 : 64;
  dim3 grid((p->w + tw - 1) / tw, (p->h + th - 1) / th, p->batch);
  cudaStream_t s = (cudaStream_t)stream;
  const float4* st = (const float4*)p->stats;
#define DCS_SA_T(TI, TO, TH, TW) \
  spat_apply_kernel<TI, TO, TH, TW><<<grid, 256, 0, s>>>((const TI*)p->x, p->chan_gate, st, p->w7, (TO*)p->y, (float2*)p->gate_out, p->h, p->w, p->channels)
#define DCS_SA(TI, TO)                                          \
  do {                                                          \
    if (th == 16 && tw == 64) DCS_SA_T(TI, TO, 16, 64);         \
    else if (th == 16) DCS_SA_T(TI, TO, 16, 16);                \
    else if (tw == 64) DCS_SA_T(TI, TO, 4, 64);                 \
    else DCS_SA_T(TI, TO, 4, 16);                               \
  } while (0)
  DCS_REQUIRE(is_dtype(p->in_dtype) && is_dtype(p->out_dtype), "dcs_spat_apply: bad dtype");
  DCS_REQUIRE(!is_h16(p->in_dtype) || !is_h16(p->out_dtype) || p->in_dtype == p->out_dtype, "dcs_spat_apply: mixed 16-bit types");
  if (p->in_dtype == DCS_F32 && p->out_dtype == DCS_F32) DCS_SA(float, float);
  else if (p->in_dtype == DCS_F32 && p->out_dtype == DCS_BF16) DCS_SA(float, __nv_bfloat16);
  else if (p->in_dtype == DCS_F32 && p->out_dtype == DCS_F16) DCS_SA(float, __half);
  else if (p->in_dtype == DCS_BF16 && p->out_dtype == DCS_F32) DCS_SA(__nv_bfloat16, float);
  else if (p->in_dtype == DCS_F16 && p->out_dtype == DCS_F32) DCS_SA(__half, float);
  else if (p->in_dtype == DCS_F16) DCS_SA(__half, __half);
  else DCS_SA(__nv_bfloat16, __nv_bfloat16);
#undef DCS_SA_T
#undef DCS_SA
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_attention_fused(const dcs_attention_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->sums && p->w1_r && p->w1_i && p->w2_r && p->w2_i && p->w7, "dcs_attention_fused: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->batch <= 65535 && p->h > 0 && p->w > 0, "dcs_attention_fused: bad shape");
  DCS_REQUIRE(pow2(p->channels) && p->channels >= 4 && p->channels <= 256 && p->reduced > 0 && p->reduced <= 16,
              "dcs_attention_fused: channels must be a power of two in [4, 256], reduced <= 16");
  const int C = p->channels;
  DCS_REQUIRE(is_dtype(p->in_dtype) && is_dtype(p->out_dtype), "dcs_attention_fused: bad dtype");
  DCS_REQUIRE(!is_h16(p->in_dtype) || !is_h16(p->out_dtype) || p->in_dtype == p->out_dtype, "dcs_attention_fused: mixed 16-bit types");
  const size_t esz = is_h16(p->in_dtype) ? 2 : 4;
  // tile: full image height when it is short (no vertical halo), else 16 rows; the widest power-of-two TW whose
  // x tile + halo stays under ~92 KB (two CTAs per SM)
  const int TH = p->h <= 32 ? p->h : 16;
  DCS_REQUIRE(TH <= 2 || TH % 4 == 0, "dcs_attention_fused: image height must be 1, 2 or a multiple of 4 (got %d)", p->h);
  int TW = 64;
  auto bytes = [&](int tw) {
    const size_t ph = TH + 2 * kFaR, pw = tw + 2 * kFaR;
    return ph * pw * C * 2 * esz + ph * pw * sizeof(float4) + (size_t)TH * tw * sizeof(float2) + (size_t)(2 * C + 16) * sizeof(float2) + 49 * sizeof(float4);
  };
  while (TW > 4 && (bytes(TW) > 100 * 1024 || TH * TW > 1024)) TW >>= 1;
  const size_t smem = bytes(TW);
  DCS_REQUIRE(smem <= 227 * 1024, "dcs_attention_fused: tile does not fit shared memory (C=%d H=%d)", C, p->h);
  FusedAttArgs a;
  a.x = p->x; a.y = p->y; a.sums = reinterpret_cast<const long long*>(p->sums); a.inv_hw = 1.f / ((float)p->h * (float)p->w);
  a.w1_r = p->w1_r; a.w1_i = p->w1_i; a.w2_r = p->w2_r; a.w2_i = p->w2_i; a.w7 = p->w7;
  a.H = p->h; a.W = p->w; a.C = C; a.clog2 = __builtin_ctz(C); a.R = p->reduced; a.TH = TH; a.TW = TW;
  dim3 grid((p->w + TW - 1) / TW, (p->h + TH - 1) / TH, p->batch);
  cudaStream_t s = (cudaStream_t)stream;
#define DCS_FA(TI, TO)                                                                                                   \
  do {                                                                                                                   \
    DCS_CUDA(cudaFuncSetAttribute(attention_fused_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attention_fused_kernel<TI, TO><<<grid, kFaThreads, smem, s>>>(a);                                                    \
  } while (0)
  if (p->in_dtype == DCS_F32 && p->out_dtype == DCS_F32) DCS_FA(float, float);
  else if (p->in_dtype == DCS_F32 && p->out_dtype == DCS_BF16) DCS_FA(float, __nv_bfloat16);
  else if (p->in_dtype == DCS_F32 && p->out_dtype == DCS_F16) DCS_FA(float, __half);
  else if (p->in_dtype == DCS_BF16 && p->out_dtype == DCS_F32) DCS_FA(__nv_bfloat16, float);
  else if (p->in_dtype == DCS_F16 && p->out_dtype == DCS_F32) DCS_FA(__half, float);
  else if (p->in_dtype == DCS_F16) DCS_FA(__half, __half);
  else DCS_FA(__nv_bfloat16, __nv_bfloat16);
#undef DCS_FA
  DCS_LAUNCHED();
  return 0;
}
