// stft.cu — STFT front-end and iSTFT back-end as shared-memory FFT kernels (no cuFFT).
//
// Replaces torch.stft at /root/reference/data.py:112-134 (config.py:72-77) and the polar round trip +
// mag_phase_2_wave + torch.istft at network_functions.py:398-401,140-150.
//
// Design (B200): a 512-point real FFT is one 256-point complex FFT (even/odd packing) computed by a HALF-WARP
// as a 16x16 four-step transform: every lane runs a fully unrolled 16-point FFT in registers, one padded
// shared-memory transpose, a second 16-point FFT.  A CTA of 256 threads = 16 half-warps transforms a chunk of
// 16 consecutive frames per iteration; results go through a [256 bins][16 frames] staging tile so that global
// traffic along the frame axis (the contiguous axis of the (B,F,T) layout) is issued as full 128-byte rows.
// The iSTFT overlap-adds 16-frame chunks into a 1024-sample shared ring in ascending frame order
// (deterministic), so the (B,T,512) frame matrix never exists in HBM.
//
// Algorithmic HBM bytes: STFT 4*L + 8*256*T per utterance; iSTFT 8*256*T + 4*32*(T-1).
#include <algorithm>
#include <stdint.h>
#include "common.cuh"

namespace dcs {

constexpr int kNfft = 512;
constexpr int kHop = 32;
constexpr int kBins = 256;        // bins 1..256 kept by the reference
constexpr int kChunk = 16;        // frames per CTA iteration (one per half-warp)
constexpr int kThreads = 256;
constexpr int kTrStride = 17;     // padded row of the 16x16 transpose tile (float2 units)
constexpr int kStStride = 17;     // padded row of the [bin][frame] staging tile (float2 units)
// One buffer per half-warp (= per frame of a chunk) serves, in turn, as the 16 x 17 transpose tile of its FFT and as the
// frame's 256 spectrum / 512 time samples: 273 float2 (odd multiple of 2 words -> the 16 frames start in different banks).
// Shared memory per CTA drops from 80 / 112 KB to 45 KB, i.e. from 2 to 4-5 resident CTAs per SM.
constexpr int kBufPitch = 273;

__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// 4-point DFT in place; INV selects e^{+i..}
template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  float2 apc = f2add(a, c), amc = f2sub(a, c), bpd = f2add(b, d), bmd = f2sub(b, d);
  a = f2add(apc, bpd);
  c = f2sub(apc, bpd);
  if (!INV) {
    b = make_float2(amc.x + bmd.y, amc.y - bmd.x);
    d = make_float2(amc.x - bmd.y, amc.y + bmd.x);
  } else {
    b = make_float2(amc.x - bmd.y, amc.y + bmd.x);
    d = make_float2(amc.x + bmd.y, amc.y - bmd.x);
  }
}

// 16-point DFT, natural order in and out, fully unrolled (4x4 four-step in registers).
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
  // cos / sin of 2*pi*m/16 for m = 0..9
  constexpr float CS[10] = {1.f, C1, C2, S1, 0.f, -S1, -C2, -C1, -1.f, -C1};
  constexpr float SN[10] = {0.f, S1, C2, C1, 1.f, C1, C2, S1, 0.f, -S1};
#pragma unroll
  for (int l = 0; l < 4; ++l) dft4<INV>(v[l], v[4 + l], v[8 + l], v[12 + l]);
  // now v[4*k1 + l] = A_l[k1]; twiddle by w16^(l*k1)
#pragma unroll
  for (int k1 = 1; k1 < 4; ++k1) {
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      const float c = CS[l * k1], s = INV ? SN[l * k1] : -SN[l * k1];
      float2 t = v[4 * k1 + l];
      v[4 * k1 + l] = make_float2(t.x * c - t.y * s, t.x * s + t.y * c);
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // v[4*k1 + k2] = Z[k1 + 4*k2]  ->  transpose to natural order
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      float2 t = v[4 * a + b];
      v[4 * a + b] = v[4 * b + a];
      v[4 * b + a] = t;
    }
  }
}

// 256-point complex DFT by one half-warp.  In: lane l holds v[r] = z[16*r + l].  Out: v[k2] = Z[l + 16*k2].
// tr: this half-warp's private 16 x kTrStride float2 tile.  tw256[k1 * 16 + l] = (cos, sin)(2*pi*l*k1/256): the inter-stage
// twiddles stored in the order the lanes read them (consecutive lanes -> consecutive words; the natural table indexed by
// (l * k1) & 255 put the 16 lanes 2 / 4 / 8 words apart for even k1: 2- to 8-way bank conflicts, 3.7 M per launch in the
// r03a profile).
template <bool INV>
__device__ __forceinline__ void fft256_halfwarp(float2 (&v)[16], int l, float2* tr, const float2* tw256) {
  fft16<INV>(v);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const float2 w = tw256[k1 * 16 + l];
    const float c = w.x, s = INV ? w.y : -w.y;
    float2 t = v[k1];
    tr[k1 * kTrStride + l] = make_float2(t.x * c - t.y * s, t.x * s + t.y * c);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = tr[l * kTrStride + j];
  __syncwarp();
  fft16<INV>(v);
}

struct FftTables {
  float2 tw256[256];
  float2 tw512[256];
  float win[kNfft];
};

__device__ __forceinline__ void init_tables(FftTables* t) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    float s, c;
    sincospif((float)(((i & 15) * (i >> 4)) & 255) / 128.f, &s, &c);   // entry [k1 = i >> 4][l = i & 15] = w256^(l k1)
    t->tw256[i] = make_float2(c, s);
    sincospif((float)i / 256.f, &s, &c);
    t->tw512[i] = make_float2(c, s);
  }
  for (int i = threadIdx.x; i < kNfft; i += blockDim.x) t->win[i] = 0.5f - 0.5f * cospif((float)i / 256.f);
}

struct StftSmem {
  FftTables tab;
  float x[(kChunk - 1) * kHop + kNfft];             // 992 samples feeding 16 frames
  float2 buf[kChunk][kBufPitch];                    // per frame: transpose tile, then Z[0..255]
};

// ADJ = true: the ADJOINT of the reference's iSTFT (mag_phase_2_wave's torch.istft, network_functions.py:140-150), i.e. the
// first backward kernel of the training step (the loss is SI-SNR on waveforms, so every gradient enters through the iSTFT):
// `audio` is the waveform gradient g (B, 32 (T-1)); it is zero-padded by 256 instead of reflected, divided by the
// overlap-add envelope sum_t w^2, and the rfft bins 0..255 are kept (the forward put spectrogram row k on rfft bin k), scaled
// by 2 for bins 1..255 and by 1 with the imaginary part dropped for bin 0 (oracle/train_oracle.istft_adjoint).
template <typename TBN, bool ADJ = false>
__global__ void __launch_bounds__(kThreads, 4) stft_kernel(const float* __restrict__ audio, float2* __restrict__ spec,
                                                        int L, int T, int chunks_per_cta,
                                                        const float* __restrict__ bn_affine, TBN* __restrict__ bn_out, int bn_real) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_raw);
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int h = tid >> 4, l = tid & 15;
  const int n_chunks = (T + kChunk - 1) / kChunk;
  const int c_begin = blockIdx.x * chunks_per_cta;
  const int c_end = min(c_begin + chunks_per_cta, n_chunks);
  init_tables(&sm.tab);
  const float* a = audio + (int64_t)b * L;
  float A00 = 1.f, A01 = 0.f, A10 = 0.f, A11 = 1.f, o0 = 0.f, o1 = 0.f;
  if (bn_affine) { A00 = bn_affine[0]; A01 = bn_affine[1]; A10 = bn_affine[2]; A11 = bn_affine[3]; o0 = bn_affine[4]; o1 = bn_affine[5]; }
  const float scale = 0.044194173824159216f;  // 1/sqrt(512)  (normalized=True)

  for (int c = c_begin; c < c_end; ++c) {
    const int t0 = c * kChunk;
    __syncthreads();  // previous iteration's readers of x/stage are done (also orders init_tables)
    // centre=True reflect padding: sample index s = 32*t + n - 256
    for (int i = tid; i < (kChunk - 1) * kHop + kNfft; i += kThreads) {
      int s = t0 * kHop + i - kNfft / 2;
      if constexpr (ADJ) {
        float v = 0.f;
        if (s >= 0 && s < L) {
          const int np = s + kNfft / 2;                      // index in the zero-padded signal
          const int t_hi = min(T - 1, np / kHop), t_lo = max(0, (np - kNfft + kHop) / kHop);
          float env = 0.f;
          for (int t = t_lo; t <= t_hi; ++t) { const float w = 0.5f - 0.5f * cospif((float)(np - kHop * t) / 256.f); env += w * w; }
          v = a[s] / env;
        }
        sm.x[i] = v;
        continue;
      }
      if (s < 0) s = -s;
      if (s >= L) s = 2 * (L - 1) - s;
      s = max(0, min(s, L - 1));  // frames past T are computed on clamped garbage and never stored
      sm.x[i] = a[s];
    }
    __syncthreads();
    {
      float2 v[16];
      const float* xf = sm.x + h * kHop;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int n = 16 * r + l;  // z[n] = (x[2n] w[2n], x[2n+1] w[2n+1])
        const float2 xs = *reinterpret_cast<const float2*>(xf + 2 * n);
        const float2 ws = *reinterpret_cast<const float2*>(sm.tab.win + 2 * n);
        v[r] = make_float2(xs.x * ws.x, xs.y * ws.y);
      }
      fft256_halfwarp<false>(v, l, sm.buf[h], sm.tab.tw256);
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) sm.buf[h][l + 16 * k2] = v[k2];   // (every lane has left the transpose tile: __syncwarp inside)
    }
    __syncthreads();
    // real-FFT post-processing + transposed store: thread (f = tid&15, k rows tid>>4 + 16*i)
    {
      const int f = tid & 15, t = t0 + f;
      if (t < T) {
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int k = (ADJ ? 0 : 1) + (tid >> 4) + 16 * i;  // output bin 1..256 (adjoint: 0..255)
          float2 X;
          if (ADJ && k == 0) {
            const float2 z0 = sm.buf[f][0];
            X = make_float2(z0.x + z0.y, 0.f);               // DC bin; its imaginary part does not reach the C2R transform
          } else if (k < 256) {
            const float2 zk = sm.buf[f][k], zm = sm.buf[f][256 - k];
            const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));  // (zk - conj zm)/(2j)
            const float2 w = sm.tab.tw512[k];  // e^{-i th} = (cos, -sin)
            X = make_float2(E.x + O.x * w.x + O.y * w.y, E.y + O.y * w.x - O.x * w.y);
          } else {
            const float2 z0 = sm.buf[f][0];
            X = make_float2(z0.x - z0.y, 0.f);
          }
          X.x *= scale; X.y *= scale;
          if (ADJ && k > 0) { X.x *= 2.f; X.y *= 2.f; }      // both Hermitian halves of bin k
          const int64_t o = ((int64_t)b * kBins + (ADJ ? k : k - 1)) * T + t;
          spec[o] = X;
          if (bn_out) {
            const float2 Xb = bn_real ? make_float2(hypotf(X.x, X.y), 0.f) : X;   // real path: BatchNorm2d of |Y| (torch.abs = hypot)
            Elem<TBN>::stc(bn_out, o, make_float2(A00 * Xb.x + A01 * Xb.y + o0, A10 * Xb.x + A11 * Xb.y + o1));
          }
        }
      }
    }
  }
}

// chunks needed so that every output sample n' < 256 + 32 (T-1) is finalised (a chunk finalises 512 samples);
// when T % 16 == 0 this adds one frame-less flush chunk.
__host__ __device__ inline int istft_chunks(int T) { return (T + 7 + kChunk - 1) / kChunk; }

struct IstftSmem {
  FftTables tab;
  float2 buf[kChunk][kBufPitch];   // per frame: X[0..255] -> transpose tile -> 512 windowed time samples
  float acc[2 * kNfft];
  float env32[kHop];               // overlap-add envelope of an interior sample (depends on n mod 32 only)
};

// polar round trip of network_functions.py:398-401 + 140-142: (|s| cos(th), |s| sin(th)), th = atan2(im, re+eps)
__device__ __forceinline__ float2 polar_roundtrip(float2 s, float eps, bool exact) {
  const float mag = sqrtf(s.x * s.x + s.y * s.y);  // torch.abs(complex64) = hypot; values here are O(1)
  const float xr = s.x + eps;
  if (exact) {
    const float th = atan2f(s.y, xr);
    return make_float2(mag * cosf(th), mag * sinf(th));
  }
  const float hy = sqrtf(xr * xr + s.y * s.y);
  if (hy == 0.f) return make_float2(mag, 0.f);
  const float inv = mag / hy;
  return make_float2(xr * inv, s.y * inv);
}

// the same round trip with bare MUFU.RSQ instead of IEEE sqrt / division (exact_polar = 2, the tensor-core mode's choice):
// mag / hypot = mag2 * rsqrt(mag2) * rsqrt(hy2), rel. error ~2e-7; ~12 instead of ~40 instructions per spectrogram value
__device__ __forceinline__ float2 polar_roundtrip_mufu(float2 s, float eps) {
  const float mag2 = s.x * s.x + s.y * s.y;
  const float xr = s.x + eps;
  const float hy2 = xr * xr + s.y * s.y;
  float r1, r2;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(fmaxf(mag2, 1e-37f)));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r2) : "f"(hy2));
  const float mag = mag2 * r1;
  const bool z = hy2 == 0.f;
  const float inv = mag * r2;
  return make_float2(z ? mag : xr * inv, z ? 0.f : s.y * inv);
}

__global__ void __launch_bounds__(kThreads, 4) istft_kernel(const float2* __restrict__ spec, const float* __restrict__ mag,
                                                         const float* __restrict__ phase, float* __restrict__ audio,
                                                         int T, int chunks_per_cta, float eps, int exact) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IstftSmem& sm = *reinterpret_cast<IstftSmem*>(smem_raw);
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int h = tid >> 4, l = tid & 15;
  const int n_chunks = istft_chunks(T);
  const int c_begin = blockIdx.x * chunks_per_cta;
  const int c_end = min(c_begin + chunks_per_cta, n_chunks);
  const int Lout = kHop * (T - 1);
  init_tables(&sm.tab);
  for (int i = tid; i < 2 * kNfft; i += kThreads) sm.acc[i] = 0.f;
  if (tid < kHop) {   // sum over the 16 frames covering a sample of hann^2 (recomputed: the table is not published yet)
    float e = 0.f;
    for (int j = 0; j < kNfft / kHop; ++j) { const float w = 0.5f - 0.5f * cospif((float)(tid + kHop * j) / 256.f); e += w * w; }
    sm.env32[tid] = e;
  }
  // unnormalised 256-pt inverse DFT with the 1/2 of the even/odd split already applied: remaining 1/256 of the
  // irfft, times sqrt(512) for normalized=True  ->  sqrt(512)/256
  const float scale = 0.08838834764831845f;
  const float2* sp = spec ? spec + (int64_t)b * kBins * T : nullptr;
  const float* mg = mag ? mag + (int64_t)b * kBins * T : nullptr;
  const float* phs = phase ? phase + (int64_t)b * kBins * T : nullptr;

  const int f_ld = tid & 15, k_ld = tid >> 4;
  for (int c = max(c_begin - 1, 0); c < c_end; ++c) {
    const bool emit = c >= c_begin;
    const int t0 = c * kChunk;
    __syncthreads();
    {  // stage [256 rfft rows][16 frames]; row k of the iSTFT input is spectrogram row k (zero row appended at 256)
      const int f = f_ld, t = t0 + f;
      if (sp) {
        // 16 independent loads per thread (rows k_ld + 16 i of frame t), then the polar round trip; three resident CTAs
        // per SM hide the HBM round trip (the register prefetch of the 2-CTA version cost 32 registers)
        float2 nx[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) nx[i] = t < T ? __ldg(sp + (int64_t)(k_ld + 16 * i) * T + t) : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          sm.buf[f][k_ld + 16 * i] = t >= T ? make_float2(0.f, 0.f)
                                     : exact == 3 ? nx[i]                              // already mag * e^{j phase} (real-path tail)
                                     : exact == 2 ? polar_roundtrip_mufu(nx[i], eps) : polar_roundtrip(nx[i], eps, exact != 0);
      } else {
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int k = k_ld + 16 * i;
          float2 s = make_float2(0.f, 0.f);
          if (t < T) { const float m_ = __ldg(mg + (int64_t)k * T + t), p_ = __ldg(phs + (int64_t)k * T + t); s = make_float2(m_ * cosf(p_), m_ * sinf(p_)); }
          sm.buf[f][k] = s;
        }
      }
    }
    __syncthreads();
    {
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int k = 16 * r + l;
        float2 xk = sm.buf[h][k];
        float2 xm;
        if (k == 0) { xk.y = 0.f; xm = make_float2(0.f, 0.f); }  // C2R ignores Im X[0]; X[256] is the zero pad row
        else { xm = sm.buf[h][256 - k]; xm.y = -xm.y; }
        const float2 E = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y + xm.y));
        const float2 D = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y - xm.y));
        const float2 w = sm.tab.tw512[k];  // e^{+i th}
        const float2 O = make_float2(D.x * w.x - D.y * w.y, D.x * w.y + D.y * w.x);
        v[r] = make_float2(E.x - O.y, E.y + O.x);  // E + j O
      }
      __syncwarp();   // every lane has read its spectrum values: the buffer becomes the transpose tile
      fft256_halfwarp<true>(v, l, sm.buf[h], sm.tab.tw256);
      float* fr = reinterpret_cast<float*>(sm.buf[h]);
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int m = l + 16 * k2;
        const float2 ws = *reinterpret_cast<const float2*>(sm.tab.win + 2 * m);
        *reinterpret_cast<float2*>(fr + 2 * m) = make_float2(v[k2].x * ws.x * scale, v[k2].y * ws.y * scale);
      }
    }
    __syncthreads();
    // overlap-add in ascending frame order into the ring, then finalise the first 512 samples of this chunk
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int off = tid + 256 * q;  // sample n' = 512*c + off
      if (off < (kChunk - 1) * kHop + kNfft) {
        float s = 0.f;
        const int h_hi = min(off / kHop, kChunk - 1);
        const int h_lo = max(0, (off - kNfft + kHop) / kHop);
        for (int hh = h_lo; hh <= h_hi; ++hh) s += reinterpret_cast<const float*>(sm.buf[hh])[off - kHop * hh];
        const int ring = (c * kNfft + off) & (2 * kNfft - 1);
        const float a = sm.acc[ring] + s;
        if (q < 2) {
          const int np = c * kNfft + off;
          const int n = np - kNfft / 2;
          if (emit && n >= 0 && n < Lout) {
            const int t_hi = min(T - 1, np / kHop);
            const int t_lo = max(0, (np - kNfft + kHop) / kHop);
            float env;
            if (t_hi - t_lo == kNfft / kHop - 1) {
              env = sm.env32[np & (kHop - 1)];      // interior sample: all 16 covering frames exist
            } else {
              env = 0.f;
              for (int t = t_lo; t <= t_hi; ++t) { const float w = sm.tab.win[np - kHop * t]; env += w * w; }
            }
            audio[(int64_t)b * Lout + n] = a / env;
          }
          sm.acc[ring] = 0.f;
        } else {
          sm.acc[ring] = a;
        }
      }
    }
  }
}

}  // namespace dcs

using namespace dcs;

extern "C" int dcs_stft_fwd(const dcs_stft_params* p, void* stream) {
  DCS_REQUIRE(p && p->audio && p->spec, "dcs_stft_fwd: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->length >= kNfft / 2 + 1, "dcs_stft_fwd: bad batch/length (%d, %d)", p->batch, p->length);
  DCS_REQUIRE(p->n_frames == p->length / kHop + 1, "dcs_stft_fwd: n_frames must be length/32+1 (got %d for L=%d)", p->n_frames, p->length);
  DCS_REQUIRE(!p->bn_out || p->bn_affine, "dcs_stft_fwd: bn_out without bn_affine");
  const int n_chunks = (p->n_frames + kChunk - 1) / kChunk;
  // chunks per CTA: the value that minimises (waves of 148 SMs x 4 resident CTAs) x (chunks per CTA)
  int cpc = 1;
  {
    const int64_t slots = 4 * (int64_t)num_sms();
    int64_t best = INT64_MAX;
    for (int c = 1; c <= 32; ++c) {
      const int64_t ctas = (int64_t)((n_chunks + c - 1) / c) * p->batch;
      const int64_t cost = ((ctas + slots - 1) / slots) * (c + 1);   // + 1: table set-up and pipeline fill of a CTA
      if (cost < best) { best = cost; cpc = c; }
    }
  }
  dim3 grid((n_chunks + cpc - 1) / cpc, p->batch);
  const size_t smem = sizeof(StftSmem);
  cudaStream_t s = (cudaStream_t)stream;
  if (p->bn_out && p->bn_dtype == DCS_BF16) {
    DCS_CUDA(cudaFuncSetAttribute(stft_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_kernel<__nv_bfloat16><<<grid, kThreads, smem, s>>>(p->audio, (float2*)p->spec, p->length, p->n_frames, cpc,
                                                            p->bn_affine, (__nv_bfloat16*)p->bn_out, p->bn_real);
  } else if (p->bn_out && p->bn_dtype == DCS_F16) {
    DCS_CUDA(cudaFuncSetAttribute(stft_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_kernel<__half><<<grid, kThreads, smem, s>>>(p->audio, (float2*)p->spec, p->length, p->n_frames, cpc,
                                                     p->bn_affine, (__half*)p->bn_out, p->bn_real);
  } else {
    DCS_CUDA(cudaFuncSetAttribute(stft_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_kernel<float><<<grid, kThreads, smem, s>>>(p->audio, (float2*)p->spec, p->length, p->n_frames, cpc,
                                                    p->bn_affine, (float*)p->bn_out, p->bn_real);
  }
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_istft_adjoint(const float* grad_audio, float* grad_spec, int batch, int n_frames, void* stream) {
  DCS_REQUIRE(grad_audio && grad_spec && batch > 0 && n_frames >= 2, "dcs_istft_adjoint: bad arguments");
  const int n_chunks = (n_frames + kChunk - 1) / kChunk;
  int cpc = 1;
  {
    const int64_t slots = 4 * (int64_t)num_sms();
    int64_t best = INT64_MAX;
    for (int c = 1; c <= std::min(32, n_chunks); ++c) {
      const int64_t ctas = (int64_t)((n_chunks + c - 1) / c) * batch;
      const int64_t cost = ((ctas + slots - 1) / slots) * (c + 1);
      if (cost < best) { best = cost; cpc = c; }
    }
  }
  dim3 grid((n_chunks + cpc - 1) / cpc, batch);
  const size_t smem = sizeof(StftSmem);
  DCS_CUDA(cudaFuncSetAttribute(stft_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stft_kernel<float, true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(grad_audio, (float2*)grad_spec, kHop * (n_frames - 1), n_frames, cpc,
                                                                           nullptr, nullptr, 0);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_istft_fwd(const dcs_istft_params* p, void* stream) {
  DCS_REQUIRE(p && p->audio && (p->spec || (p->mag && p->phase)), "dcs_istft_fwd: null pointer");
  DCS_REQUIRE(p->batch > 0 && p->n_frames >= 2, "dcs_istft_fwd: bad batch/n_frames (%d, %d)", p->batch, p->n_frames);
  const int n_chunks = istft_chunks(p->n_frames);
  // chunks per CTA: every CTA re-computes one halo chunk, so the cost of a choice is (waves of 148 SMs x 4 resident
  // CTAs) x (chunks per CTA + 1)
  int cpc = std::min(8, n_chunks);
  {
    const int64_t slots = 4 * (int64_t)num_sms();
    int64_t best = INT64_MAX;
    for (int c = 1; c <= std::min(32, n_chunks); ++c) {
      const int64_t ctas = (int64_t)((n_chunks + c - 1) / c) * p->batch;
      const int64_t cost = ((ctas + slots - 1) / slots) * (c + 1);
      if (cost < best) { best = cost; cpc = c; }
    }
  }
  dim3 grid((n_chunks + cpc - 1) / cpc, p->batch);
  const size_t smem = sizeof(IstftSmem);
  DCS_CUDA(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  istft_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>((const float2*)p->spec, p->mag, p->phase, p->audio,
                                                               p->n_frames, cpc, p->atan2_eps, p->exact_polar);
  DCS_LAUNCHED();
  return 0;
}
