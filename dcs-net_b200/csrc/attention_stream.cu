// attention_stream.cu — ComplexChannelAttention + ComplexSpatialAttention as ONE streaming pass (bf16 / tensor-core mode).
//
// Reference: c_network.py:53-84 (modules), 208-211 / 219-220 (y = SA(u) * u, u = CA(x) * x), pools
// network_functions.py:114-138, ComplexSigmoid 107-112.
//
// The three-kernel form (spat_stats -> spat_apply, attention.cu) reads x twice, round-trips 16 B / pixel of statistics
// through HBM and is issue-bound: at C = 8 the 7x7 gate conv alone is 196 packed FMAs per pixel (r01r profile: 283 us
// for a 262 MB tensor, 69 % issue-active).  Here a CTA owns a column strip of TW pixels of one image and walks DOWN the
// rows:
//   * x rows arrive by 1-D bulk copies (cp.async.bulk, one per row segment + halo: channels-last rows are contiguous)
//     into a ring of NR rows, NR - 4 rows ahead of their use: x is read from HBM once (halo columns are L2 hits);
//   * row r: per-pixel statistics (mean_c u, max_c Re u, max_c Im u) -> one shared-memory row, tf32-rounded;
//   * the 7x7 complex gate conv runs on the tensor cores as mma.sync.m16n8k8 TF32: for ONE statistics row the
//     A operand is the row itself read as a sliding matrix A[m][kx*4 + ci] = row[(m + kx) * 4 + ci] (16 output pixels
//     x 7 taps x 4 real inputs, no im2col), B holds the taps with N = (ky, re/im): 8 MMAs per 16 pixels give the 14
//     row-partials P[ky][re/im].  Partial ky of statistics row r belongs to output row r + 3 - ky: lane t of a quad
//     owns ky = t and ky = t + 4, so it keeps a 4-deep register ring of pending output rows (the ring slot is the MMA's
//     C operand 4 rows later); a quad reduction assembles the finished row.  3.1 warp instructions per pixel instead
//     of ~8 for the packed-FMA form, and no 7-row statistics window in shared memory;
//   * output row r - 3: gate_s * gate_c * x from the ring -> HBM (16-byte stores).
// HBM traffic: x once in, y once out.  fp32 mode keeps the exact CUDA-core kernels (TF32 products would break 1e-5).
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <type_traits>
#include <algorithm>

namespace dcs {

constexpr int kAsThreads = 256, kAsRing = 8;

struct AttStreamArgs {
  const void* x; void* y;           // 16-bit storage (fp16 or bf16: the kernel's template parameter)
  const long long* sums; float inv_hw;
  const float *w1_r, *w1_i, *w2_r, *w2_i, *w7;
  int H, W, R, NR;
};

// round-to-nearest for a tf32 MMA operand: the tensor core ignores the low 13 mantissa bits, so adding half an ulp of
// tf32 to the bit pattern is all that is needed (finite inputs)
__device__ __forceinline__ float tf32_rna(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }
__device__ __forceinline__ float sigmoid_ex2(float v) {   // MUFU.EX2 + MUFU.RCP, abs. error ~3e-7
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// packed fp32x2 arithmetic (FMUL2 / FADD2 / FFMA2): the kernel is issue-bound, two lanes of work per slot
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
      "l"(*reinterpret_cast<const unsigned long long*>(&b)), "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
// a 16-byte vector = 4 complex 16-bit values (e0 e1 e2 e3) as two element PAIRS in split form: re = (e_a.re, e_b.re), im
// likewise.  Eight single-result conversions either way (bf16: shift / mask, fp16: the two halves of HADD2.F32).
struct CPair { float2 re, im; };
template <typename T>
__device__ __forceinline__ void unpack_pairs(const uint4 q, CPair& p01, CPair& p23) {
  const float2 e0 = unpack_h2<T>(q.x), e1 = unpack_h2<T>(q.y), e2 = unpack_h2<T>(q.z), e3 = unpack_h2<T>(q.w);
  p01.re = make_float2(e0.x, e1.x); p01.im = make_float2(e0.y, e1.y);
  p23.re = make_float2(e2.x, e3.x); p23.im = make_float2(e2.y, e3.y);
}
// per-thread channel gate of an element pair: (g.re, g.im, -g.im) as pairs
struct GPair { float2 re, im, nim; };
__device__ __forceinline__ CPair cmul_pair(const GPair& g, const CPair& v) {
  CPair u;
  u.re = fma2(g.re, v.re, mul2(g.nim, v.im));
  u.im = fma2(g.re, v.im, mul2(g.im, v.re));
  return u;
}

// shared-memory access by 32-bit address (no generic-address arithmetic in the row loop)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts64f(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

// C channels (power of two >= 8), strip width TW (multiple of 16, <= 128).
// REAL: the real network's CBAM (r_network.py:8-40) on a pair tensor — C pairs = 2C real channels, element-wise gates,
// max-pool channel gate, 2 statistics per pixel, one real spatial gate (same data movement, same MMA structure with the
// imaginary rows / inputs zero).
template <typename T, int C, int TW, bool REAL = false>
__global__ void __launch_bounds__(kAsThreads, 2) attention_stream_kernel(const AttStreamArgs a) {
  constexpr int PW = TW + 6;                                  // strip + 3-pixel halo each side
  constexpr int VPP = C / 4;                                  // 16-byte vectors per pixel
  constexpr int GMAX = kAsThreads / PW;                       // lanes per pixel in the statistics pass: largest power of two
  constexpr int G = GMAX >= 8 ? (VPP >= 8 ? 8 : VPP) : GMAX >= 4 ? (VPP >= 4 ? 4 : VPP) : GMAX >= 2 ? 2 : 1;
  constexpr int VPL = VPP / G;                                // vectors per lane (statistics)
  constexpr int NSEG = TW / 16;                               // 16-pixel MMA segments (one warp each)
  constexpr bool OWN = NSEG >= 7;                             // a warp applies the 16 pixels whose gate it computed (TW = 112, 128)
  constexpr int NAPP = OWN ? VPP / 2 : (TW * VPP + kAsThreads - 1) / kAsThreads;   // vectors per thread (product)
  constexpr uint32_t ROW_BYTES = (uint32_t)PW * C * 4;
  constexpr int ST_PITCH = (PW + 2) * 4;                      // floats per statistics row (2 zero pixels of padding)
  static_assert(G >= 1 && VPL >= 1 && VPL * G == VPP && PW * G <= kAsThreads, "statistics mapping");
  static_assert(OWN || (kAsThreads % VPP == 0), "product mapping");

  extern __shared__ __align__(128) unsigned char as_smem[];
  const int NR = a.NR, H = a.H, W = a.W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, x0 = blockIdx.x * TW;
  unsigned char* xs = as_smem;                                       // [NR][PW][C] bf16 complex
  float* st = reinterpret_cast<float*>(xs + (size_t)NR * ROW_BYTES);  // [4][ST_PITCH]
  float* sg = st + 4 * ST_PITCH;                                     // [2 rows][2][TW] spatial gates (re plane, im plane) of the step's rows
  float2* gs = reinterpret_cast<float2*>(sg + 4 * TW);               // [C] channel gate
  float2* avg = gs + C;                                              // [C]
  float2* hid = avg + C;                                             // [16]
  uint64_t* full = reinterpret_cast<uint64_t*>(hid + 16);            // [kAsRing]
  const uint32_t xs_u32 = smem_u32(xs), full_u32 = smem_u32(full);

  const int xa = max(x0 - 3, 0), xe = min(x0 + TW + 3, W);           // image columns this strip reads
  const uint32_t seg_bytes = (uint32_t)(xe - xa) * C * 4;
  const uint32_t seg_off = (uint32_t)(xa - (x0 - 3)) * C * 4;
  const T* xsrc = reinterpret_cast<const T*>(a.x) + ((int64_t)b * H * W + xa) * C * 2;
  auto issue_row = [&](int r) {                                      // one thread; ring slot r & 7 (r < NR when H < 8)
    const uint32_t bar = full_u32 + 8 * (r & (kAsRing - 1));
    mbar_expect_tx(bar, seg_bytes);
    bulk_g2s(xs_u32 + (r & (kAsRing - 1)) * ROW_BYTES + seg_off, xsrc + (int64_t)r * W * C * 2, seg_bytes, bar);
  };
  if (tid == 0) {
    for (int i = 0; i < kAsRing; ++i) mbar_init(full_u32 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * ST_PITCH; i += kAsThreads) st[i] = 0.f;
  for (int c = tid; c < C; c += kAsThreads) {
    if constexpr (REAL) avg[c] = make_float2(pool_max_value(a.sums, ((int64_t)b * C + c) * 2), pool_max_value(a.sums, ((int64_t)b * C + c) * 2 + 1));
    else avg[c] = make_float2(pool_mean(a.sums, ((int64_t)b * C + c) * 2, a.inv_hw), pool_mean(a.sums, ((int64_t)b * C + c) * 2 + 1, a.inv_hw));
  }
  __syncthreads();
  if (tid == 0)
    for (int r = 0; r < min(NR, H); ++r) issue_row(r);

  // ---- channel gate (ComplexChannelAttention): sigmoid_c(2 W2 crelu(W1 avg)), recomputed per CTA (C * R complex MACs)
  //      REAL (RealChannelAttention, r_network.py:20-25): sigmoid(W2 relu(W1 max)) over the 2C real channels; w1 (R, 2C), w2 (2C, R)
  for (int r = warp; r < a.R; r += kAsThreads / 32) {
    float re = 0.f, im = 0.f;
    for (int c = lane; c < C; c += 32) {
      if constexpr (REAL) {
        re += a.w1_r[r * 2 * C + 2 * c] * avg[c].x + a.w1_r[r * 2 * C + 2 * c + 1] * avg[c].y;
      } else {
        const float wr = a.w1_r[r * C + c], wi = a.w1_i[r * C + c];
        re += wr * avg[c].x - wi * avg[c].y;
        im += wr * avg[c].y + wi * avg[c].x;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
  }
  __syncthreads();
  for (int c = tid; c < C; c += kAsThreads) {
    float re = 0.f, im = 0.f;
    for (int r = 0; r < a.R; ++r) {
      if constexpr (REAL) {
        re += a.w2_r[(2 * c) * a.R + r] * hid[r].x;
        im += a.w2_r[(2 * c + 1) * a.R + r] * hid[r].x;
      } else {
        const float wr = a.w2_r[c * a.R + r], wi = a.w2_i[c * a.R + r];
        re += wr * hid[r].x - wi * hid[r].y;
        im += wr * hid[r].y + wi * hid[r].x;
      }
    }
    gs[c] = REAL ? make_float2(sigmoidf_(re), sigmoidf_(im)) : make_float2(sigmoidf_(2.f * re), sigmoidf_(2.f * im));
  }

  // ---- gate conv operands (ComplexConv2d(2, 1, 7, padding=3, bias=False) as a real 4 -> 2 conv).  Per statistics row the
  //      MMA is  D[m][n] = sum_kk Wm[m][kk] * S[n][kk]:  n = 8 pixels, kk = kx * 4 + ci (28 of 32 used),
  //      ci = (mean.re, mean.im, max.re, max.im), and m = 16 rows of row-partials: lane group g owns rows g and g + 8 with
  //      q = g >> 1, output part o = g & 1:   row g: ky = q (q < 3; zero for q = 3),  row g + 8: ky = q + 4 (q < 3), ky = 3 (q = 3).
  //      The taps are the A operand (constant registers); the statistics row is the B operand, whose fragment
  //      (k = t, t + 4 of k-step s  <->  kk = 8 t + 2 s, 8 t + 2 s + 1) is 8 consecutive floats of the row: two LDS.128.
  float afr[4][4];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = 8 * t + 2 * s + (r >> 1), kx = kk >> 2, ci = kk & 3;
      const int q = g >> 1, o = g & 1;
      const int ky = (r & 1) == 0 ? (q < 3 ? q : 7) : (q < 3 ? 4 + q : 3);
      float v = 0.f;
      if (kx < 7 && ky < 7) {
        if constexpr (REAL) {   // Conv2d(2, 1, 7): inputs ci = (mean, max, -, -), one real output (the o = 0 rows)
          if (o == 0 && ci < 2) v = a.w7[ci * 49 + ky * 7 + kx];
        } else {
          const int idx = (ci >> 1) * 49 + ky * 7 + kx;
          const float wr = a.w7[idx], wi = a.w7[98 + idx];
          v = o == 0 ? ((ci & 1) ? -wi : wr) : ((ci & 1) ? wr : wi);
        }
      }
      afr[s][r] = tf32_rna(v);
    }
  __syncthreads();

  // ---- per-thread constants of the three phases (the pixel / channels a thread works on never change with the row)
  const uint32_t st_u32 = smem_u32(st), sg_u32 = smem_u32(sg);
  // statistics: pixel sp of the padded row, vectors sub + k G
  const int sp = tid / G, sub = tid % G;
  const bool s_act = sp < PW;
  const bool s_in = s_act && (unsigned)(x0 - 3 + sp) < (unsigned)W;
  const uint32_t s_ld = xs_u32 + (uint32_t)(s_act ? sp : 0) * C * 4 + sub * 16;
  const uint32_t s_st = st_u32 + (uint32_t)(s_act ? sp : 0) * 16;
  GPair sgate[VPL][2];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const float2 ga = gs[(sub + k * G) * 4 + 2 * h2], gb = gs[(sub + k * G) * 4 + 2 * h2 + 1];
      sgate[k][h2].re = make_float2(ga.x, gb.x); sgate[k][h2].im = make_float2(ga.y, gb.y);
      sgate[k][h2].nim = make_float2(-ga.y, -gb.y);
    }
  // product: vectors i = i0 + k * istep of the strip row -> pixel i / VPP, vector i % VPP (constant per thread)
  const int i0 = OWN ? warp * 16 * VPP + lane : tid;
  constexpr int istep = OWN ? 32 : kAsThreads;
  GPair agate[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const float2 ga = gs[(i0 % VPP) * 4 + 2 * h2], gb = gs[(i0 % VPP) * 4 + 2 * h2 + 1];
    agate[h2].re = make_float2(ga.x, gb.x); agate[h2].im = make_float2(ga.y, gb.y); agate[h2].nim = make_float2(-ga.y, -gb.y);
  }
  bool a_ok[NAPP];
#pragma unroll
  for (int k = 0; k < NAPP; ++k) a_ok[k] = i0 + k * istep < TW * VPP && x0 + (i0 + k * istep) / VPP < W;
  const uint32_t a_ld = xs_u32 + 3 * C * 4 + (uint32_t)i0 * 16;          // + ring slot * ROW_BYTES + k * istep * 16
  const uint32_t a_sg = sg_u32 + (uint32_t)(i0 / VPP) * 4;               // plane re; + TW * 4 plane im; + k * (istep / VPP) * 4
  T* yp = reinterpret_cast<T*>(a.y) + ((int64_t)b * H * W + x0) * C * 2 + (int64_t)i0 * 8;   // output row 0; + y_pitch per row
  const int64_t y_pitch = (int64_t)W * C * 2;
  const float invC = 1.f / (float)C;
  // conv: this lane's B values in a statistics row (float4 index warp * 16 + g + 2 t; second pixel group + 8), gate
  // store of the lanes that end up with the row totals (q = 2): plane o, pixels warp * 16 + 2 t (+ 8)
  const int q = g >> 1;
  const uint32_t c_ld = st_u32 + (uint32_t)(warp * 16 + g + 2 * t) * 16;
  const uint32_t c_st = sg_u32 + (uint32_t)((g & 1) * TW + warp * 16 + 2 * t) * 4;
  const int src_lane = (lane & 7) | (((q + 3) & 3) << 3);

  float slot[4][2][2], run[2][2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      run[h2][e] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) slot[i][h2][e] = 0.f;
    }
  // Two rows per step: the statistics, MMA chains and products of the two rows are independent instruction streams (the
  // kernel is bound by issue latency at 2 CTAs per SM), one CTA barrier per row pair.
  auto stats_row = [&](int row, uint32_t sbuf) {
    const uint32_t ring = (uint32_t)(row & (kAsRing - 1));
    float2 sre = make_float2(0.f, 0.f), sim = make_float2(0.f, 0.f);
    float mr = -INFINITY, mi = -INFINITY;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {   // lanes outside the image read (valid) shared memory and are masked at the store
      CPair p01, p23;
      unpack_pairs<T>(lds128(s_ld + ring * ROW_BYTES + k * G * 16), p01, p23);
      CPair u01, u23;
      if constexpr (REAL) {   // element-wise gates: the (re, im) slots are independent real channels
        u01.re = mul2(sgate[k][0].re, p01.re); u01.im = mul2(sgate[k][0].im, p01.im);
        u23.re = mul2(sgate[k][1].re, p23.re); u23.im = mul2(sgate[k][1].im, p23.im);
      } else {
        u01 = cmul_pair(sgate[k][0], p01); u23 = cmul_pair(sgate[k][1], p23);
      }
      sre = add2(sre, add2(u01.re, u23.re));
      sim = add2(sim, add2(u01.im, u23.im));
      mr = fmaxf(fmaxf(mr, fmaxf(u01.re.x, u01.re.y)), fmaxf(u23.re.x, u23.re.y));
      mi = fmaxf(fmaxf(mi, fmaxf(u01.im.x, u01.im.y)), fmaxf(u23.im.x, u23.im.y));
    }
    float sr = sre.x + sre.y, si = sim.x + sim.y;
    if constexpr (REAL) { sr = 0.5f * (sr + si); si = 0.f; mr = fmaxf(mr, mi); mi = 0.f; }   // mean / max over all 2C real channels
    if (G > 1) {   // every lane takes part
#pragma unroll
      for (int o = G >> 1; o; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o);
        mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
      }
    }
    if (s_act && sub == 0) {
      if (s_in) {
        if constexpr (REAL) sts128f(s_st + sbuf, tf32_rna(sr * invC), tf32_rna(mr), 0.f, 0.f);   // ci = (mean, max, -, -)
        else sts128f(s_st + sbuf, tf32_rna(sr * invC), tf32_rna(si * invC), tf32_rna(mr), tf32_rna(mi));
      }
      else sts128f(s_st + sbuf, 0.f, 0.f, 0.f, 0.f);
    }
  };
  auto apply_row = [&](int yo, uint32_t gbuf) {
    const uint32_t ring = (uint32_t)(yo & (kAsRing - 1));
#pragma unroll
    for (int k = 0; k < NAPP; ++k) {
      if (a_ok[k]) {
        CPair p01, p23;
        unpack_pairs<T>(lds128(a_ld + ring * ROW_BYTES + k * istep * 16), p01, p23);
        float2 gsp;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gsp.x) : "r"(a_sg + gbuf + k * (istep / VPP) * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gsp.y) : "r"(a_sg + gbuf + TW * 4 + k * (istep / VPP) * 4));
        const float2 gre = make_float2(gsp.x, gsp.x), gim = make_float2(gsp.y, gsp.y), gnim = make_float2(-gsp.y, -gsp.y);
        float2 r01, i01, r23, i23;
        if constexpr (REAL) {   // y = gate_s * gate_c * x, everything real
          r01 = mul2(gre, mul2(agate[0].re, p01.re)); i01 = mul2(gre, mul2(agate[0].im, p01.im));
          r23 = mul2(gre, mul2(agate[1].re, p23.re)); i23 = mul2(gre, mul2(agate[1].im, p23.im));
        } else {
          const CPair u01 = cmul_pair(agate[0], p01), u23 = cmul_pair(agate[1], p23);
          r01 = fma2(gre, u01.re, mul2(gnim, u01.im)); i01 = fma2(gre, u01.im, mul2(gim, u01.re));
          r23 = fma2(gre, u23.re, mul2(gnim, u23.im)); i23 = fma2(gre, u23.im, mul2(gim, u23.re));
        }
        *reinterpret_cast<uint4*>(yp + (int64_t)yo * y_pitch + (size_t)k * istep * 8) =
            make_uint4(pack_h2<T>(r01.x, i01.x), pack_h2<T>(r01.y, i01.y), pack_h2<T>(r23.x, i23.x), pack_h2<T>(r23.y, i23.y));
      }
    }
  };
  constexpr uint32_t kStBuf = ST_PITCH * 4, kSgBuf = 2 * TW * 4;
  const int nsteps = H + 3;
  int next_row = min(NR, H);                       // thread 0: rows below this are issued
  for (int rr0 = 0; rr0 < nsteps; rr0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; u += 2) {
      const int rr = rr0 + u;                      // this step: statistics rows rr, rr + 1 -> output rows rr - 3, rr - 2
      if (rr >= nsteps) break;
      const uint32_t stb = u * kStBuf;             // statistics buffers (u, u + 1): a fast warp's next step must not touch them
      // ---- 1. statistics (zero rows outside the image = the gate conv's zero padding)
      if (rr + 1 < H) {
        mbar_wait(full_u32 + 8 * (rr & (kAsRing - 1)), (uint32_t)((rr >> 3) & 1));
        mbar_wait(full_u32 + 8 * ((rr + 1) & (kAsRing - 1)), (uint32_t)(((rr + 1) >> 3) & 1));
        stats_row(rr, stb);
        stats_row(rr + 1, stb + kStBuf);
      } else {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (rr + v < H) {
            mbar_wait(full_u32 + 8 * ((rr + v) & (kAsRing - 1)), (uint32_t)(((rr + v) >> 3) & 1));
            stats_row(rr + v, stb + v * kStBuf);
          } else if (tid < PW) {
            sts128f(st_u32 + stb + v * kStBuf + tid * 16, 0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      __syncthreads();
      // rows up to rr - 4 have been multiplied out (previous step): their ring slots take rows up to rr + 4
      if (tid == 0)
        while (next_row < H && next_row <= rr + 4) issue_row(next_row++);

      // ---- 2. gate conv: a statistics row's 14 row-partials on the tensor cores.  Partial ky of statistics row r belongs
      //         to output row r + 3 - ky.  A lane with q < 3 owns ky = q (accumulator rows g: opens a pending output row
      //         in slot[r & 3]) and ky = q + 4 (rows g + 8, four statistics rows later: the SAME slot, fed back as the C
      //         operand, closes it); q = 3 owns ky = 3 in rows g + 8.  So after the MMAs every lane holds its finished
      //         share of output row r - 1 - q (q = 3: row r), and the row total travels q = 3 -> 0 -> 1 -> 2, one hop
      //         per statistics row: the q = 2 lanes end with the complete output row r - 3.
      if (warp < NSEG) {
        float d[2][2][4];
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 v0 = lds128f(c_ld + stb + v * kStBuf + h2 * 128), v1 = lds128f(c_ld + stb + v * kStBuf + h2 * 128 + 16);
            d[v][h2][0] = 0.f; d[v][h2][1] = 0.f; d[v][h2][2] = slot[u + v][h2][0]; d[v][h2][3] = slot[u + v][h2][1];
            mma_tf32_16x8x8(d[v][h2], afr[0], v0.x, v0.y);
            mma_tf32_16x8x8(d[v][h2], afr[1], v0.z, v0.w);
            mma_tf32_16x8x8(d[v][h2], afr[2], v1.x, v1.y);
            mma_tf32_16x8x8(d[v][h2], afr[3], v1.z, v1.w);
            slot[u + v][h2][0] = d[v][h2][0]; slot[u + v][h2][1] = d[v][h2][1];
          }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float in = __shfl_sync(0xffffffffu, run[h2][e], src_lane);   // the previous lane's total one row earlier
              run[h2][e] = q == 3 ? d[v][h2][2 + e] : d[v][h2][2 + e] + in;
            }
          if (q == 2) {   // gate of pixels 2 t, 2 t + 1 (+ 8) of this warp's segment, part o, output row rr + v - 3
            sts64f(c_st + v * kSgBuf, sigmoid_ex2(run[0][0]), sigmoid_ex2(run[0][1]));
            sts64f(c_st + v * kSgBuf + 32, sigmoid_ex2(run[1][0]), sigmoid_ex2(run[1][1]));
          }
        }
      }
      if (OWN) __syncwarp(); else __syncthreads();

      // ---- 3. y = gate_s * (gate_c * x) for output rows rr - 3, rr - 2
      if (rr >= 3 && rr - 2 < H) {
        apply_row(rr - 3, 0);
        apply_row(rr - 2, kSgBuf);
      } else {
        if (rr >= 3 && rr - 3 < H) apply_row(rr - 3, 0);
        if (rr >= 2 && rr - 2 < H) apply_row(rr - 2, kSgBuf);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Whole-strip ("tile") kernel for the SHORT, WIDE-CHANNEL tensors (C >= 32 pairs: H * C <= 1024, i.e. encoder[2..6] outputs
// and decoder[0..3] outputs: 32 x 250 x 32 ... 2 x 250 x 128 per image).  The row-streaming kernel above walks down the rows
// with three CTA barriers per row pair and ONE warp on the gate conv of a 16-pixel strip; on these tensors (2 ... 32 rows)
// it is bound by the fill / drain latency of that pipeline (58-70 us for a 65 MB tensor = 1 TB/s).  Here a full-height column
// strip of 16 pixels (+3 halo each side) is ONE shared-memory tile (H * 22 * C * 4 B <= 88 KB, 2 CTAs per SM), every row
// arrives by its own bulk copy, and the CTA runs three flat phases separated by two barriers:
//   1. statistics of all H x 22 pixels (threads own a fixed vector subset of a pixel; conflict-free rotated reads),
//      written as fp16 (mean.re, mean.im, max.re, max.im) — the same 11-bit significand as the streaming kernel's tf32;
//   2. the 7x7 gate conv of FOUR output rows x 16 pixels per warp as 20 mma.sync.m16n8k16 (fp16, fp32 accumulate):
//      M = 16 pixels, N = (output row j = 0..3, re / im), K = (statistics row y' = 0..9, kx = 0..7, ci = 0..3); the A
//      fragment is the statistics tile itself read as a sliding window (A[m][(y', kx, ci)] = st[r0 + y'][m + kx][ci], four
//      conflict-free LDS.32 per MMA, no im2col), B = the taps shifted by the output row, W[y' - j][kx][ci][o];
//   3. y = gate_s * (gate_c * x) for the H x 16 strip pixels, 16-byte loads / stores, consecutive lanes on consecutive
//      addresses.
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}


// Generalised to row BANDS for the tall few-channel tensors (C = 8: 128 x 1000 pixels per image): a tile is RH output rows x
// TW output columns plus a 3-pixel halo on every side that lies inside the image (statistics of the halo pixels are
// recomputed: 1.34 x the pixels at RH = 32, TW = 48), grid = (column strips, images, bands).  The row-streaming kernel needs
// ~14 warp instructions per pixel at C = 8 (register ring of pending conv rows, three CTA barriers per row pair); the flat
// phases need ~8.
template <int C, bool REAL, int NT, int TW>
__global__ void __launch_bounds__(NT, 2) attention_tile_kernel(const AttStreamArgs a) {
  using T = __half;
  constexpr int PW = TW + 6;                  // x-tile row: strip + 3-pixel halo each side
  constexpr int SP = TW + 8;                  // statistics row pitch in pixels (the MMA's sliding window reads up to pixel TW + 6)
  constexpr int NSEG = TW / 16;               // 16-pixel MMA segments per row
  constexpr int VPP = C / 4;                  // 16-byte vectors per pixel
  constexpr int VPL = VPP < 1024 / NT ? VPP : 1024 / NT;   // vectors per lane in the statistics phase (<= 4 at 256 threads)
  constexpr int G = VPP / VPL;                // lanes per pixel in the statistics phase
  constexpr int PPI = NT / G;                 // pixels per statistics iteration (a multiple of 8: the read rotation is per thread)
  constexpr int RSH = VPP >= 8 ? 0 : (VPP == 4 ? 1 : 2);   // rotation = (pixel >> RSH) & (VPL - 1): conflict-free quarter warps
  constexpr int QS = NT / VPP;                // output pixels per product iteration
  constexpr int NW7 = REAL ? 98 : 196;
  static_assert(TW % 16 == 0 && VPP * 4 == C && VPL * G == VPP && (VPL == 2 || VPL == 4) && G <= 16, "tile attention geometry");

  extern __shared__ __align__(128) unsigned char as_smem[];
  const int H = a.H, W = a.W, RH = a.NR;      // RH = band height (= H for the whole-strip launches)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, x0 = blockIdx.x * TW, r0 = blockIdx.z * RH;
  const int rh = min(RH, H - r0);                                  // output rows of this band
  const int lo = max(r0 - 3, 0), hi = min(r0 + rh + 3, H);         // image rows held in the x tile
  const int nxr = hi - lo, soff = lo - r0 + 3;                     // statistics slot of x-tile row i = i + soff (slot s <-> image row r0 - 3 + s)
  const int NXR = min(RH + 6, H), RHP = (RH + 3) & ~3;             // allocation sizes (same for every band)
  unsigned char* xs = as_smem;                                                         // [NXR][PW][C] complex fp16
  uint2* st = reinterpret_cast<uint2*>(xs + (size_t)NXR * PW * C * 4);                  // [RHP + 6][SP] 4 x fp16 statistics
  float2* sg = reinterpret_cast<float2*>(st + (size_t)(RHP + 6) * SP);                  // [RHP][TW] spatial gate
  float2* gs = sg + RHP * TW;                                                          // [C] channel gate
  float2* avg = gs + C;                                                                // [C]
  float2* hid = avg + C;                                                               // [16]
  float* w7s = reinterpret_cast<float*>(hid + 16);                                     // [196]
  uint2* bt = reinterpret_cast<uint2*>(w7s + 196);                                     // [20][32] B fragments of the gate conv
  uint64_t* full = reinterpret_cast<uint64_t*>(bt + 20 * 32);                          // [4] one per group of RG rows
  const uint32_t xs_u32 = smem_u32(xs), st_u32 = smem_u32(st), full_u32 = smem_u32(full);
  const int RG = (nxr + 3) >> 2;                                                       // rows per arrival group (<= 4 groups)

  // ---- rows -> shared memory (one bulk copy per row: the strip's columns are contiguous in the channels-last layout)
  const int xa = max(x0 - 3, 0), xe = min(x0 + TW + 3, W);
  const uint32_t seg_bytes = (uint32_t)(xe - xa) * C * 4;
  const uint32_t seg_off = (uint32_t)(xa - (x0 - 3)) * C * 4;
  if (tid < 4) {
    mbar_init(full_u32 + 8 * tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < nxr) {
    const int grp = tid / RG;
    const uint32_t bar = full_u32 + 8 * grp;
    if (tid == grp * RG) mbar_expect_tx(bar, (uint32_t)(min(RG, nxr - grp * RG)) * seg_bytes);   // one arrival per group; the copies may land in any order
    bulk_g2s(xs_u32 + (uint32_t)tid * (PW * C * 4) + seg_off,
             reinterpret_cast<const T*>(a.x) + (((int64_t)b * H + lo + tid) * W + xa) * C * 2, seg_bytes, bar);
  }
  for (int i = tid; i < (RHP + 6) * SP; i += NT) st[i] = make_uint2(0u, 0u);
  for (int i = tid; i < NW7; i += NT) w7s[i] = a.w7[i];
  for (int c = tid; c < C; c += NT) {
    if constexpr (REAL) avg[c] = make_float2(pool_max_value(a.sums, ((int64_t)b * C + c) * 2), pool_max_value(a.sums, ((int64_t)b * C + c) * 2 + 1));
    else avg[c] = make_float2(pool_mean(a.sums, ((int64_t)b * C + c) * 2, a.inv_hw), pool_mean(a.sums, ((int64_t)b * C + c) * 2 + 1, a.inv_hw));
  }
  __syncthreads();

  // ---- B fragments of the gate conv (phase 2), one (b0, b1) pair per (k-step, lane): k-step ks = (statistics row y' = ks / 2,
  //      kx half s = ks % 2); lane (g, t): column n = g = (output row j = g / 2, part o = g % 2); the k order inside a k-step is
  //      chosen so that a lane's A registers are the two WORDS of one statistics pixel (one LDS.64): k pair (2 t, 2 t + 1) =
  //      (re, im) of the mean at kx = 4 s + t, pair (2 t + 8, 2 t + 9) = (re, im) of the max at the same kx; tap row ky = y' - j.
  for (int e = tid; e < 20 * 32; e += NT) {
    const int ks = e >> 5, ln = e & 31, gg = ln >> 2, tt = ln & 3;
    const int ky = (ks >> 1) - (gg >> 1), o = gg & 1;
    uint32_t bfr[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int kx = 4 * (ks & 1) + tt, cp = h2;      // k pair (2 t, 2 t + 1) = (re, im) of the mean at kx, pair (2 t + 8, 2 t + 9) = of the max
      float v0 = 0.f, v1 = 0.f;
      if (ky >= 0 && ky < 7 && kx < 7) {
        if constexpr (REAL) {                 // Conv2d(2, 1, 7): ci = (mean, max, -, -), one real output (o = 0)
          if (o == 0 && cp == 0) { v0 = w7s[ky * 7 + kx]; v1 = w7s[49 + ky * 7 + kx]; }
        } else {
          const float wr = w7s[cp * 49 + ky * 7 + kx], wi = w7s[98 + cp * 49 + ky * 7 + kx];
          v0 = o == 0 ? wr : wi; v1 = o == 0 ? -wi : wr;
        }
      }
      bfr[h2] = pack_f16x2(v0, v1);
    }
    bt[e] = make_uint2(bfr[0], bfr[1]);
  }

  // ---- channel gate (same arithmetic as the streaming kernel; all loads of a pass in flight together)
  for (int r = warp; r < a.R; r += NT / 32) {
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int c = lane; c < C; c += 32) {
      if constexpr (REAL) {
        re += a.w1_r[r * 2 * C + 2 * c] * avg[c].x + a.w1_r[r * 2 * C + 2 * c + 1] * avg[c].y;
      } else {
        const float wr = a.w1_r[r * C + c], wi = a.w1_i[r * C + c];
        re += wr * avg[c].x - wi * avg[c].y;
        im += wr * avg[c].y + wi * avg[c].x;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
  }
  __syncthreads();
  if (tid < C) {
    const int c = tid;
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r < a.R) {
        if constexpr (REAL) {
          re += a.w2_r[(2 * c) * a.R + r] * hid[r].x;
          im += a.w2_r[(2 * c + 1) * a.R + r] * hid[r].x;
        } else {
          const float wr = a.w2_r[c * a.R + r], wi = a.w2_i[c * a.R + r];
          re += wr * hid[r].x - wi * hid[r].y;
          im += wr * hid[r].y + wi * hid[r].x;
        }
      }
    }
    gs[c] = REAL ? make_float2(sigmoidf_(re), sigmoidf_(im)) : make_float2(sigmoidf_(2.f * re), sigmoidf_(2.f * im));
  }
  __syncthreads();

  // ---- 1. statistics: thread = (pixel f0 + j PPI, vectors sub + G ((k + rot) & (VPL - 1))); the rotation spreads the lanes of
  //         a quarter warp over all 32 banks (pixels are 32 ... 512 bytes apart)
  {
    const int sub = tid % G, f0 = tid / G, rot = (f0 >> RSH) & (VPL - 1);
    GPair sgate[VPL][2];
    uint32_t voff[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = sub + G * ((k + rot) & (VPL - 1));
      voff[k] = (uint32_t)vi * 16;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const float2 ga = gs[vi * 4 + 2 * h2], gb = gs[vi * 4 + 2 * h2 + 1];
        sgate[k][h2].re = make_float2(ga.x, gb.x); sgate[k][h2].im = make_float2(ga.y, gb.y);
        sgate[k][h2].nim = make_float2(-ga.y, -gb.y);
      }
    }
    const int npix = nxr * PW;
    const float invC = 1.f / (float)C;
    int waited = -1;
    for (int fb = 0; fb < npix; fb += PPI) {
      const bool valid = fb + f0 < npix;
      const int f = valid ? fb + f0 : npix - 1;
      const int r = f / PW, p = f - r * PW;                              // x-tile row, padded column
      const int grp = (r >= RG) + (r >= 2 * RG) + (r >= 3 * RG);       // r / RG without the runtime division
      if (grp > waited) { for (int q = waited + 1; q <= grp; ++q) mbar_wait(full_u32 + 8 * q, 0); waited = grp; }
      const uint32_t base = xs_u32 + (uint32_t)f * (C * 4);
      float2 sre = make_float2(0.f, 0.f), sim = make_float2(0.f, 0.f);
      float mr = -INFINITY, mi = -INFINITY;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        CPair p01, p23, u01, u23;
        unpack_pairs<T>(lds128(base + voff[k]), p01, p23);
        if constexpr (REAL) {
          u01.re = mul2(sgate[k][0].re, p01.re); u01.im = mul2(sgate[k][0].im, p01.im);
          u23.re = mul2(sgate[k][1].re, p23.re); u23.im = mul2(sgate[k][1].im, p23.im);
        } else {
          u01 = cmul_pair(sgate[k][0], p01); u23 = cmul_pair(sgate[k][1], p23);
        }
        sre = add2(sre, add2(u01.re, u23.re));
        sim = add2(sim, add2(u01.im, u23.im));
        mr = fmaxf(fmaxf(mr, fmaxf(u01.re.x, u01.re.y)), fmaxf(u23.re.x, u23.re.y));
        mi = fmaxf(fmaxf(mi, fmaxf(u01.im.x, u01.im.y)), fmaxf(u23.im.x, u23.im.y));
      }
      float sr = sre.x + sre.y, si = sim.x + sim.y;
      if constexpr (REAL) { sr = 0.5f * (sr + si); si = 0.f; mr = fmaxf(mr, mi); mi = 0.f; }
#pragma unroll
      for (int o = G >> 1; o; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o);
        mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
      }
      if (valid && sub == 0 && (unsigned)(x0 - 3 + p) < (unsigned)W)      // columns outside the image keep their zeros
        st[(r + soff) * SP + p] = REAL ? make_uint2(pack_f16x2(sr * invC, mr), 0u)       // ci = (mean, max, -, -)
                                       : make_uint2(pack_f16x2(sr * invC, si * invC), pack_f16x2(mr, mi));
    }
  }
  __syncthreads();

  // ---- 2. gate conv: unit = (group of 4 output rows, 16-pixel segment), two accumulators (half the dependent MMA chain)
  {
    const int nunits = ((rh + 3) >> 2) * NSEG;
    const uint32_t b_base = smem_u32(bt) + (uint32_t)lane * 8;
    for (int u = warp; u < nunits; u += NT / 32) {
      const int rg = u / NSEG, sgm = u - rg * NSEG;
      const int rr0 = 4 * rg;
      float d[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t a_base = st_u32 + (uint32_t)((rr0 * SP + sgm * 16 + g + t) * 8);
#pragma unroll
      for (int ks = 0; ks < 20; ++ks) {
        const uint32_t aa = a_base + (uint32_t)(((ks >> 1) * SP + 4 * (ks & 1)) * 8);
        const uint2 bb = lds64u(b_base + ks * 256), alo = lds64u(aa), ahi = lds64u(aa + 64);   // (a0, a2) of pixel g + kx, (a1, a3) of pixel g + 8 + kx
        if (ks & 1) mma_f16_16x8x16(d2, alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
        else mma_f16_16x8x16(d, alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] += d2[i];
      const int rl = rr0 + t;                   // accumulator columns 2 t, 2 t + 1 = (re, im) of output row rr0 + t
      if (rl < rh) {
        float2* o = sg + rl * TW + sgm * 16 + g;
        if constexpr (REAL) {
          o[0] = make_float2(sigmoid_ex2(d[0]), 0.f);
          o[8] = make_float2(sigmoid_ex2(d[2]), 0.f);
        } else {
          o[0] = make_float2(sigmoid_ex2(d[0]), sigmoid_ex2(d[1]));
          o[8] = make_float2(sigmoid_ex2(d[2]), sigmoid_ex2(d[3]));
        }
      }
    }
  }
  __syncthreads();

  // ---- 3. y = gate_s * (gate_c * x)
  {
    const int vi = tid % VPP, q0 = tid / VPP;
    GPair agate[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const float2 ga = gs[vi * 4 + 2 * h2], gb = gs[vi * 4 + 2 * h2 + 1];
      agate[h2].re = make_float2(ga.x, gb.x); agate[h2].im = make_float2(ga.y, gb.y); agate[h2].nim = make_float2(-ga.y, -gb.y);
    }
    const int nq = rh * TW;
    const int xrow0 = r0 - lo;                  // x-tile row of output row 0 of the band
    T* ybase = reinterpret_cast<T*>(a.y) + ((((int64_t)b * H + r0) * W + x0) * C + vi * 4) * 2;
#pragma unroll 4
    for (int q = q0; q < nq; q += QS) {
      const int rl = q / TW, px = q - rl * TW;
      if (x0 + px < W) {
        CPair p01, p23;
        unpack_pairs<T>(lds128(xs_u32 + (uint32_t)((((xrow0 + rl) * PW + 3 + px) * VPP + vi) * 16)), p01, p23);
        const float2 gsp = sg[q];
        const float2 gre = make_float2(gsp.x, gsp.x), gim = make_float2(gsp.y, gsp.y), gnim = make_float2(-gsp.y, -gsp.y);
        float2 r01, i01, r23, i23;
        if constexpr (REAL) {
          r01 = mul2(gre, mul2(agate[0].re, p01.re)); i01 = mul2(gre, mul2(agate[0].im, p01.im));
          r23 = mul2(gre, mul2(agate[1].re, p23.re)); i23 = mul2(gre, mul2(agate[1].im, p23.im));
        } else {
          const CPair u01 = cmul_pair(agate[0], p01), u23 = cmul_pair(agate[1], p23);
          r01 = fma2(gre, u01.re, mul2(gnim, u01.im)); i01 = fma2(gre, u01.im, mul2(gim, u01.re));
          r23 = fma2(gre, u23.re, mul2(gnim, u23.im)); i23 = fma2(gre, u23.im, mul2(gim, u23.re));
        }
        *reinterpret_cast<uint4*>(ybase + ((int64_t)rl * W + px) * C * 2) =
            make_uint4(pack_h2<T>(r01.x, i01.x), pack_h2<T>(r01.y, i01.y), pack_h2<T>(r23.x, i23.x), pack_h2<T>(r23.y, i23.y));
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Row-group streaming kernel for the TALL few-channel tensors (C = 8, 16): the flat phases of the tile kernel above, but a CTA
// walks DOWN a column strip in groups of FOUR rows — x rows in a ring of NG groups (bulk copies, one barrier per group, two
// groups ahead), statistics rows in a 16-row ring — so nothing is recomputed vertically (halo: 6 columns per strip) and the
// gate prologue is paid once per strip.  Step s: statistics of rows 4 s .. 4 s + 3; barrier; gate conv of output rows
// 4 s - 4 .. 4 s - 1 (one 16-pixel segment per warp, 20 fp16 MMAs from statistics rows 4 s - 7 .. 4 s + 2); barrier; product of
// those rows.  Two CTA barriers per four rows (the older streaming kernel: three per two rows, plus a register ring of pending
// conv rows and one shuffle hop per row).
template <int C, bool REAL, int TW, int NG, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) attention_rows_kernel(const AttStreamArgs a) {
  using T = __half;
  constexpr int NT = 256;
  constexpr int PW = TW + 6, SP = TW + 8, NSEG = TW / 16;
  constexpr int VPP = C / 4;
  constexpr int VPL = VPP < 4 ? VPP : 4;
  constexpr int G = VPP / VPL;
  constexpr int PPI = NT / G;
  constexpr int RSH = VPP >= 8 ? 0 : (VPP == 4 ? 1 : 2);
  constexpr int QS = NT / VPP;
  constexpr int NW7 = REAL ? 98 : 196;
  constexpr int NXR = 4 * NG;                 // x ring rows
  constexpr uint32_t ROWB = (uint32_t)PW * C * 4;
  static_assert(TW % 16 == 0 && NSEG <= 8 && NG == 4 && VPL * G == VPP, "row-group attention geometry");

  extern __shared__ __align__(128) unsigned char as_smem[];
  const int H = a.H, W = a.W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, x0 = blockIdx.x * TW;
  unsigned char* xs = as_smem;                                                         // [NXR][PW][C] complex fp16
  uint2* st = reinterpret_cast<uint2*>(xs + (size_t)NXR * ROWB);                        // [16][SP] 4 x fp16 statistics (row & 15)
  float2* sg = reinterpret_cast<float2*>(st + (size_t)16 * SP);                         // [2][4][TW] spatial gates of two row groups
  float2* gs = sg + 8 * TW;                                                            // [C] channel gate
  float2* avg = gs + C;
  float2* hid = avg + C;
  float* w7s = reinterpret_cast<float*>(hid + 16);
  uint2* bt = reinterpret_cast<uint2*>(w7s + 196);                                     // [20][32] B fragments
  uint64_t* full = reinterpret_cast<uint64_t*>(bt + 20 * 32);                          // [NG]
  const uint32_t xs_u32 = smem_u32(xs), st_u32 = smem_u32(st), full_u32 = smem_u32(full);

  const int xa = max(x0 - 3, 0), xe = min(x0 + TW + 3, W);
  const uint32_t seg_bytes = (uint32_t)(xe - xa) * C * 4;
  const uint32_t seg_off = (uint32_t)(xa - (x0 - 3)) * C * 4;
  const T* xsrc = reinterpret_cast<const T*>(a.x) + ((int64_t)b * H * W + xa) * C * 2;
  // rows 4 grp .. 4 grp + 3 -> ring group grp % NG (threads 0..3, one row each; thread 0 arms the barrier)
  auto issue_group = [&](int grp) {
    const int r = 4 * grp + tid;
    const int n = min(4, H - 4 * grp);
    const uint32_t bar = full_u32 + 8 * (grp % NG);
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)n * seg_bytes);
    if (tid < n) bulk_g2s(xs_u32 + (uint32_t)((grp % NG) * 4 + tid) * ROWB + seg_off, xsrc + (int64_t)r * W * C * 2, seg_bytes, bar);
  };
  if (tid < NG) {
    mbar_init(full_u32 + 8 * tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 4) issue_group(0);                          // (group s + 1 is issued after step s's barrier)
  for (int i = tid; i < 16 * SP; i += NT) st[i] = make_uint2(0u, 0u);
  for (int i = tid; i < NW7; i += NT) w7s[i] = a.w7[i];
  for (int c = tid; c < C; c += NT) {
    if constexpr (REAL) avg[c] = make_float2(pool_max_value(a.sums, ((int64_t)b * C + c) * 2), pool_max_value(a.sums, ((int64_t)b * C + c) * 2 + 1));
    else avg[c] = make_float2(pool_mean(a.sums, ((int64_t)b * C + c) * 2, a.inv_hw), pool_mean(a.sums, ((int64_t)b * C + c) * 2 + 1, a.inv_hw));
  }
  __syncthreads();
  for (int e = tid; e < 20 * 32; e += NT) {            // B fragments of the gate conv (see attention_tile_kernel)
    const int ks = e >> 5, ln = e & 31, gg = ln >> 2, tt = ln & 3;
    const int ky = (ks >> 1) - (gg >> 1), o = gg & 1;
    uint32_t bfr[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int kx = 4 * (ks & 1) + tt, cp = h2;      // k pair (2 t, 2 t + 1) = (re, im) of the mean at kx, pair (2 t + 8, 2 t + 9) = of the max
      float v0 = 0.f, v1 = 0.f;
      if (ky >= 0 && ky < 7 && kx < 7) {
        if constexpr (REAL) {
          if (o == 0 && cp == 0) { v0 = w7s[ky * 7 + kx]; v1 = w7s[49 + ky * 7 + kx]; }
        } else {
          const float wr = w7s[cp * 49 + ky * 7 + kx], wi = w7s[98 + cp * 49 + ky * 7 + kx];
          v0 = o == 0 ? wr : wi; v1 = o == 0 ? -wi : wr;
        }
      }
      bfr[h2] = pack_f16x2(v0, v1);
    }
    bt[e] = make_uint2(bfr[0], bfr[1]);
  }
  for (int r = warp; r < a.R; r += NT / 32) {           // channel gate (as in the other kernels)
    float re = 0.f, im = 0.f;
    for (int c = lane; c < C; c += 32) {
      if constexpr (REAL) {
        re += a.w1_r[r * 2 * C + 2 * c] * avg[c].x + a.w1_r[r * 2 * C + 2 * c + 1] * avg[c].y;
      } else {
        const float wr = a.w1_r[r * C + c], wi = a.w1_i[r * C + c];
        re += wr * avg[c].x - wi * avg[c].y;
        im += wr * avg[c].y + wi * avg[c].x;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
  }
  __syncthreads();
  if (tid < C) {
    const int c = tid;
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r < a.R) {
        if constexpr (REAL) {
          re += a.w2_r[(2 * c) * a.R + r] * hid[r].x;
          im += a.w2_r[(2 * c + 1) * a.R + r] * hid[r].x;
        } else {
          const float wr = a.w2_r[c * a.R + r], wi = a.w2_i[c * a.R + r];
          re += wr * hid[r].x - wi * hid[r].y;
          im += wr * hid[r].y + wi * hid[r].x;
        }
      }
    }
    gs[c] = REAL ? make_float2(sigmoidf_(re), sigmoidf_(im)) : make_float2(sigmoidf_(2.f * re), sigmoidf_(2.f * im));
  }
  __syncthreads();

  // ---- per-thread constants of the phases
  const int sub = tid % G, f0 = tid / G, rot = (f0 >> RSH) & (VPL - 1);
  GPair sgate[VPL][2];
  uint32_t voff[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = sub + G * ((k + rot) & (VPL - 1));
    voff[k] = (uint32_t)vi * 16;
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const float2 ga = gs[vi * 4 + 2 * h2], gb = gs[vi * 4 + 2 * h2 + 1];
      sgate[k][h2].re = make_float2(ga.x, gb.x); sgate[k][h2].im = make_float2(ga.y, gb.y);
      sgate[k][h2].nim = make_float2(-ga.y, -gb.y);
    }
  }
  const int avi = tid % VPP, q0 = tid / VPP;
  GPair agate[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const float2 ga = gs[avi * 4 + 2 * h2], gb = gs[avi * 4 + 2 * h2 + 1];
    agate[h2].re = make_float2(ga.x, gb.x); agate[h2].im = make_float2(ga.y, gb.y); agate[h2].nim = make_float2(-ga.y, -gb.y);
  }
  const float invC = 1.f / (float)C;
  const uint32_t b_base = smem_u32(bt) + (uint32_t)lane * 8;
  T* ybase = reinterpret_cast<T*>(a.y) + (((int64_t)b * H * W + x0) * C + avi * 4) * 2;

  // Step s: statistics of group s; ONE CTA barrier; gate conv of group s - 1 (-> sg[(s - 1) & 1]) and the product of group
  // s - 2 (gates from sg[s & 1], written one step earlier) as one instruction stream per warp: the product's independent work
  // fills the MMA chain's latency, and no second barrier is needed (the hazards — statistics rows, x groups, gate buffers — are
  // all separated by the next step's barrier).
  const int ngroups = (H + 3) / 4;
  for (int s = 0; s < ngroups + 2; ++s) {
    // ---- 1. statistics of rows 4 s .. 4 s + 3 (zero rows below the image: the conv's zero padding, and the ring holds old rows)
    if (s <= ngroups) {
      const int rbase = 4 * s;
      if (rbase < H) mbar_wait(full_u32 + 8 * (s % NG), (uint32_t)((s / NG) & 1));
      for (int fb = 0; fb < 4 * PW; fb += PPI) {
        const bool valid = fb + f0 < 4 * PW;
        const int f = valid ? fb + f0 : 4 * PW - 1;
        const int r4 = f / PW, p = f - r4 * PW;
        const int row = rbase + r4;
        const uint32_t base = xs_u32 + (uint32_t)(((s % NG) * 4 + r4) * PW + p) * (C * 4);
        float2 sre = make_float2(0.f, 0.f), sim = make_float2(0.f, 0.f);
        float mr = -INFINITY, mi = -INFINITY;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          CPair p01, p23, u01, u23;
          unpack_pairs<T>(lds128(base + voff[k]), p01, p23);
          if constexpr (REAL) {
            u01.re = mul2(sgate[k][0].re, p01.re); u01.im = mul2(sgate[k][0].im, p01.im);
            u23.re = mul2(sgate[k][1].re, p23.re); u23.im = mul2(sgate[k][1].im, p23.im);
          } else {
            u01 = cmul_pair(sgate[k][0], p01); u23 = cmul_pair(sgate[k][1], p23);
          }
          sre = add2(sre, add2(u01.re, u23.re));
          sim = add2(sim, add2(u01.im, u23.im));
          mr = fmaxf(fmaxf(mr, fmaxf(u01.re.x, u01.re.y)), fmaxf(u23.re.x, u23.re.y));
          mi = fmaxf(fmaxf(mi, fmaxf(u01.im.x, u01.im.y)), fmaxf(u23.im.x, u23.im.y));
        }
        float sr = sre.x + sre.y, si = sim.x + sim.y;
        if constexpr (REAL) { sr = 0.5f * (sr + si); si = 0.f; mr = fmaxf(mr, mi); mi = 0.f; }
#pragma unroll
        for (int o = G >> 1; o; o >>= 1) {
          sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o);
          mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
        }
        if (valid && sub == 0) {
          const bool in = row < H && (unsigned)(x0 - 3 + p) < (unsigned)W;
          uint2 v = make_uint2(0u, 0u);
          if (in) v = REAL ? make_uint2(pack_f16x2(sr * invC, mr), 0u) : make_uint2(pack_f16x2(sr * invC, si * invC), pack_f16x2(mr, mi));
          st[(row & 15) * SP + p] = v;
        }
      }
    }
    __syncthreads();
    // every thread has finished the product of group s - 3 (previous step): its ring group takes the rows of group s + 1
    if (tid < 4 && 4 * (s + 1) < H) issue_group(s + 1);
    // ---- 2. gate conv of output rows 4 (s - 1) .. + 3 from statistics rows 4 s - 7 .. 4 s + 2: issue the MMAs (four chains) ...
    const bool do_conv = s >= 1 && s <= ngroups && warp < NSEG;
    float d[4][4];
    if (do_conv) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
      const uint32_t a_col = (uint32_t)((warp * 16 + g + t) * 8);
      const int rb = 4 * s - 7;
#pragma unroll
      for (int ks = 0; ks < 20; ++ks) {
        const uint32_t aa = st_u32 + (uint32_t)(((rb + (ks >> 1)) & 15) * SP + 4 * (ks & 1)) * 8 + a_col;
        const uint2 bb = lds64u(b_base + ks * 256), alo = lds64u(aa), ahi = lds64u(aa + 64);
        mma_f16_16x8x16(d[ks & 3], alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
      }
    }
    // ---- 3. ... y = gate_s * (gate_c * x) for rows 4 (s - 2) .. + 3 while they complete ...
    if (s >= 2) {
      const int ob = 4 * (s - 2);
      const int nq = min(4, H - ob) * TW;
      const uint32_t xg = xs_u32 + (uint32_t)(((s - 2) % NG) * 4) * ROWB + 3 * C * 4 + (uint32_t)avi * 16;
      const float2* sgr = sg + (s & 1) * 4 * TW;
      T* yrow = ybase + (int64_t)ob * W * C * 2;
#pragma unroll 4
      for (int q = q0; q < nq; q += QS) {
        const int rl = q / TW, px = q - rl * TW;
        if (x0 + px < W) {
          CPair p01, p23;
          unpack_pairs<T>(lds128(xg + (uint32_t)(rl * PW + px) * (C * 4)), p01, p23);
          const float2 gsp = sgr[q];
          const float2 gre = make_float2(gsp.x, gsp.x), gim = make_float2(gsp.y, gsp.y), gnim = make_float2(-gsp.y, -gsp.y);
          float2 r01, i01, r23, i23;
          if constexpr (REAL) {
            r01 = mul2(gre, mul2(agate[0].re, p01.re)); i01 = mul2(gre, mul2(agate[0].im, p01.im));
            r23 = mul2(gre, mul2(agate[1].re, p23.re)); i23 = mul2(gre, mul2(agate[1].im, p23.im));
          } else {
            const CPair u01 = cmul_pair(agate[0], p01), u23 = cmul_pair(agate[1], p23);
            r01 = fma2(gre, u01.re, mul2(gnim, u01.im)); i01 = fma2(gre, u01.im, mul2(gim, u01.re));
            r23 = fma2(gre, u23.re, mul2(gnim, u23.im)); i23 = fma2(gre, u23.im, mul2(gim, u23.re));
          }
          *reinterpret_cast<uint4*>(yrow + ((int64_t)rl * W + px) * C * 2) =
              make_uint4(pack_h2<T>(r01.x, i01.x), pack_h2<T>(r01.y, i01.y), pack_h2<T>(r23.x, i23.x), pack_h2<T>(r23.y, i23.y));
        }
      }
    }
    // ---- ... and store the finished gates of group s - 1 for the next step's product
    if (do_conv) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d[0][i] = (d[0][i] + d[1][i]) + (d[2][i] + d[3][i]);
      float2* o = sg + ((s - 1) & 1) * 4 * TW + t * TW + warp * 16 + g;         // accumulator columns 2 t, 2 t + 1 = (re, im) of row 4 (s - 1) + t
      if constexpr (REAL) {
        o[0] = make_float2(sigmoid_ex2(d[0][0]), 0.f);
        o[8] = make_float2(sigmoid_ex2(d[0][2]), 0.f);
      } else {
        o[0] = make_float2(sigmoid_ex2(d[0][0]), sigmoid_ex2(d[0][1]));
        o[8] = make_float2(sigmoid_ex2(d[0][2]), sigmoid_ex2(d[0][3]));
      }
    }
  }
}

}  // namespace dcs

using namespace dcs;

template <typename T, int C, int TW, bool REAL = false>
static int launch_attention_stream(const dcs_attention_params* p, cudaStream_t s) {
  const int NR = p->h < kAsRing ? p->h : kAsRing;
  const size_t pw = TW + 6;
  const size_t smem = (size_t)NR * pw * C * 4 + 4 * (pw + 2) * 16 + (size_t)TW * 16 + (size_t)(2 * C + 16) * 8 + kAsRing * 8;
  DCS_REQUIRE(smem <= 227 * 1024, "dcs_attention_stream: strip does not fit shared memory (C=%d)", C);
  AttStreamArgs a;
  a.x = p->x; a.y = p->y; a.sums = reinterpret_cast<const long long*>(p->sums); a.inv_hw = 1.f / ((float)p->h * (float)p->w);
  a.w1_r = p->w1_r; a.w1_i = p->w1_i; a.w2_r = p->w2_r; a.w2_i = p->w2_i; a.w7 = p->w7;
  a.H = p->h; a.W = p->w; a.R = p->reduced; a.NR = NR;
  // (set on every call: the attribute is per device, and one process may drive several GPUs)
  DCS_CUDA(cudaFuncSetAttribute(attention_stream_kernel<T, C, TW, REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p->w + TW - 1) / TW, p->batch);
  attention_stream_kernel<T, C, TW, REAL><<<grid, kAsThreads, smem, s>>>(a);
  DCS_LAUNCHED();
  return 0;
}

template <int C, bool REAL, int NT, int TW>
static int launch_attention_tile_nt(const dcs_attention_params* p, cudaStream_t s, int RH) {
  const int H = p->h;
  if (RH <= 0 || RH > H) RH = H;
  const int NXR = std::min(RH + 6, H), RHP = (RH + 3) & ~3;
  const size_t smem = (size_t)NXR * (TW + 6) * C * 4 + (size_t)(RHP + 6) * (TW + 8) * 8 + (size_t)RHP * TW * 8 + (size_t)(2 * C + 16) * 8 + 196 * 4 + 20 * 32 * 8 + 4 * 8;
  DCS_REQUIRE(smem <= 113 * 1024 && NXR <= NT, "dcs_attention_stream: tile does not fit shared memory (C=%d, H=%d, band %d)", C, H, RH);
  AttStreamArgs a;
  a.x = p->x; a.y = p->y; a.sums = reinterpret_cast<const long long*>(p->sums); a.inv_hw = 1.f / ((float)p->h * (float)p->w);
  a.w1_r = p->w1_r; a.w1_i = p->w1_i; a.w2_r = p->w2_r; a.w2_i = p->w2_i; a.w7 = p->w7;
  a.H = p->h; a.W = p->w; a.R = p->reduced; a.NR = RH;
  DCS_CUDA(cudaFuncSetAttribute(attention_tile_kernel<C, REAL, NT, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p->w + TW - 1) / TW, p->batch, (H + RH - 1) / RH);
  attention_tile_kernel<C, REAL, NT, TW><<<grid, NT, smem, s>>>(a);
  DCS_LAUNCHED();
  return 0;
}
// 8 warps (4 vectors per lane in the statistics phase, ~100 registers) by default; DCS_ATT_TILE_NT=512 selects the 16-warp
// instance (2 vectors per lane, 64 registers), measured 10 % slower on B200 (more shuffles, longer barriers)
template <int C, bool REAL>
static int launch_attention_tile(const dcs_attention_params* p, cudaStream_t s) {
  static const bool nt512 = [] { const char* e = getenv("DCS_ATT_TILE_NT"); return e && atoi(e) == 512; }();
  return nt512 ? launch_attention_tile_nt<C, REAL, 512, 16>(p, s, 0) : launch_attention_tile_nt<C, REAL, 256, 16>(p, s, 0);
}

template <int C, bool REAL, int TW, int NG, int MINB = 2>
static int launch_attention_rows(const dcs_attention_params* p, cudaStream_t s) {
  const size_t smem = (size_t)4 * NG * (TW + 6) * C * 4 + (size_t)16 * (TW + 8) * 8 + (size_t)8 * TW * 8 + (size_t)(2 * C + 16) * 8 + 196 * 4 + 20 * 32 * 8 + 4 * 8;
  DCS_REQUIRE(smem <= 113 * 1024, "dcs_attention_stream: row-group ring does not fit shared memory (C=%d, TW=%d)", C, TW);
  AttStreamArgs a;
  a.x = p->x; a.y = p->y; a.sums = reinterpret_cast<const long long*>(p->sums); a.inv_hw = 1.f / ((float)p->h * (float)p->w);
  a.w1_r = p->w1_r; a.w1_i = p->w1_i; a.w2_r = p->w2_r; a.w2_i = p->w2_i; a.w7 = p->w7;
  a.H = p->h; a.W = p->w; a.R = p->reduced; a.NR = 0;
  DCS_CUDA(cudaFuncSetAttribute(attention_rows_kernel<C, REAL, TW, NG, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p->w + TW - 1) / TW, p->batch);
  attention_rows_kernel<C, REAL, TW, NG, MINB><<<grid, 256, smem, s>>>(a);
  DCS_LAUNCHED();
  return 0;
}
// Default: C = 8 only.  Measured on B200 (batch 64 x 4 s): C = 8 tensors 207 -> 156 us (110 M -> 77 M warp instructions); at C = 16
// (64- / 80-pixel strips: the ring of 16 rows x 64 B pixels limits the strip width) 87 -> 91 us, so those keep the older kernel.
// Also tried: 80-pixel strips at three CTAs per SM (<= 80 registers): 0.95 vs 0.88 ms for the attention stage.
static int att_rows_mask() {   // DCS_ATT_ROWS: bit 0: C = 8, bit 1: C = 16 through attention_rows_kernel (0: the older streaming kernel)
  static const int m = [] { const char* e = getenv("DCS_ATT_ROWS"); return e ? atoi(e) : 1; }();
  return m;
}

// DCS_ATT_TILE=0 keeps the row-streaming kernel for every tensor, DCS_ATT_BAND=0 for the tall few-channel tensors (A/B runs)
static bool att_tile_enabled() {
  static const bool on = [] { const char* e = getenv("DCS_ATT_TILE"); return !(e && e[0] == '0'); }();
  return on;
}
static int att_band_rh() {     // band height at C = 8 (DCS_ATT_BAND_RH: 16 = three resident CTAs, 32 = two, less halo)
  static const int m = [] { const char* e = getenv("DCS_ATT_BAND_RH"); const int v = e ? atoi(e) : 32; return v >= 8 && v <= 32 ? v : 32; }();
  return m;
}
// Row bands for C = 8 / 16 are OFF by default: measured on B200 (batch 64 x 4 s) the band kernel needs 102 M warp instructions
// for a C = 8 tensor against the streaming kernel's 110 M at the same issue rate (225 vs 212 us) — the halo statistics (1.34 x),
// the per-CTA prologue (5376 CTAs) and the gate conv's fragment loads eat what the missing barriers save.
static int att_band_mask() {   // DCS_ATT_BAND: bit 0: C = 8, bit 1: C = 16
  static const int m = [] { const char* e = getenv("DCS_ATT_BAND"); return e ? atoi(e) : 0; }();
  return m;
}

template <typename T, bool REAL = false>
static int dispatch_attention_stream(const dcs_attention_params* p, cudaStream_t s) {
  if constexpr (std::is_same<T, __half>::value) {
    // short wide-channel tensors: the whole column strip is one shared-memory tile (attention_tile_kernel)
    if (att_tile_enabled() && p->channels >= 32 && p->h <= 32 && p->h * p->channels <= 1024) {
      switch (p->channels) {
        case 32: return launch_attention_tile<32, REAL>(p, s);
        case 64: return launch_attention_tile<64, REAL>(p, s);
        case 128: return launch_attention_tile<128, REAL>(p, s);
        default: break;
      }
    }
    // tall few-channel tensors: row bands of the same kernel (32 x 48 pixels at C = 8, 24 x 32 at C = 16, + halo)
    if (att_tile_enabled() && p->channels == 8 && (att_band_mask() & 1) && p->h >= 32 && p->w >= 48)
      return launch_attention_tile_nt<8, REAL, 256, 48>(p, s, att_band_rh());
    if (att_tile_enabled() && p->channels == 16 && (att_band_mask() & 2) && p->h >= 24 && p->w >= 32)
      return launch_attention_tile_nt<16, REAL, 256, 32>(p, s, 24);
  }
  if constexpr (std::is_same<T, __half>::value) {
    // tall few-channel tensors: row groups of four (attention_rows_kernel); strip width by wave fill like the older kernel
    auto fill = [&](int tw) { const int64_t ctas = (int64_t)((p->w + tw - 1) / tw) * p->batch, slots = 2 * num_sms(); return ((ctas + slots - 1) / slots) * (tw + 6); };
    if (p->channels == 8 && (att_rows_mask() & 1) && p->h >= 8 && p->w > 64)
      return fill(112) < fill(128) ? launch_attention_rows<8, REAL, 112, 4>(p, s) : launch_attention_rows<8, REAL, 128, 4>(p, s);
    if (p->channels == 16 && (att_rows_mask() & 2) && p->h >= 8 && p->w > 48)
      return fill(64) <= fill(80) ? launch_attention_rows<16, REAL, 64, 4>(p, s) : launch_attention_rows<16, REAL, 80, 4>(p, s);
  }
  // strip width: 128 pixels (every warp owns 16) for the few-channel tensors; narrow strips where a row of C channels is
  // long (ring of 8 rows) and the tensor has few pixels (enough CTAs), or where the image itself is narrow
  const int w = p->w;
  // 112- or 128-pixel strips for the wide images: whichever fills the 2-CTAs-per-SM waves better (time of a CTA ~ TW + 6)
  auto cost = [&](int tw) { const int64_t ctas = (int64_t)((w + tw - 1) / tw) * p->batch, slots = 2 * num_sms(); return ((ctas + slots - 1) / slots) * (tw + 6); };
  const bool wide112 = cost(112) < cost(128);
  switch (p->channels) {
    case 8: return w > 64 ? (wide112 ? launch_attention_stream<T, 8, 112, REAL>(p, s) : launch_attention_stream<T, 8, 128, REAL>(p, s))
                          : w > 32 ? launch_attention_stream<T, 8, 64, REAL>(p, s) : launch_attention_stream<T, 8, 32, REAL>(p, s);
    case 16: return w > 64 ? (wide112 ? launch_attention_stream<T, 16, 112, REAL>(p, s) : launch_attention_stream<T, 16, 128, REAL>(p, s))
                           : w > 32 ? launch_attention_stream<T, 16, 64, REAL>(p, s) : launch_attention_stream<T, 16, 32, REAL>(p, s);
    case 32: return w > 16 ? launch_attention_stream<T, 32, 32, REAL>(p, s) : launch_attention_stream<T, 32, 16, REAL>(p, s);
    case 64: return w > 16 ? launch_attention_stream<T, 64, 32, REAL>(p, s) : launch_attention_stream<T, 64, 16, REAL>(p, s);
    case 128: return launch_attention_stream<T, 128, 16, REAL>(p, s);
    default: break;
  }
  return set_error(-1, "dcs_attention_stream: channels must be 8, 16, 32, 64 or 128 (got %d)", p->channels);
}

extern "C" int dcs_attention_stream(const dcs_attention_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->sums && p->w1_r && p->w2_r && p->w7 && (p->real || (p->w1_i && p->w2_i)), "dcs_attention_stream: null pointer");
  DCS_REQUIRE(is_h16(p->in_dtype) && p->out_dtype == p->in_dtype,
              "dcs_attention_stream: 16-bit storage only, same type in and out (the fp32 mode uses dcs_spat_stats / dcs_spat_apply)");
  DCS_REQUIRE(p->batch > 0 && p->batch <= 65535 && p->h > 0 && p->w > 0, "dcs_attention_stream: bad shape");
  DCS_REQUIRE(p->reduced > 0 && p->reduced <= 16, "dcs_attention_stream: reduced must be in [1, 16]");
  cudaStream_t s = (cudaStream_t)stream;
  if (p->real) {
    DCS_REQUIRE(p->in_dtype == DCS_F16, "dcs_attention_stream: the real variant is built for fp16 storage (bf16: dcs_real_attention_fwd)");
    return dispatch_attention_stream<__half, true>(p, s);
  }
  return p->in_dtype == DCS_F16 ? dispatch_attention_stream<__half>(p, s) : dispatch_attention_stream<__nv_bfloat16>(p, s);
}
