// train.cu — first kernels of the TRAINING step (BASELINE configs[4]): the stages that have no inference
// twin.  Contracts (closed forms, each checked against autograd) are stated in oracle/train_oracle.py; reference lines:
//   train-mode ComplexBatchNorm2d ...... complexPyTorch 0.3 complexLayers.py; c_network.py:101,113,148
//   loss (SI-SNR on waveforms) .......... network_functions.py:30-42 (SiSNR), 168-208 (calc_loss, types 6 / 0)
//   mask tail adjoint ................... network_functions.py:236-247 (dcs) / 271-275 (dc), 77-88 (bound_cRM), 398-401 (polar)
//   decoder up-sampling + concat adjoint. c_network.py:214-215
// (the iSTFT adjoint is the STFT kernel in ADJ mode, stft.cu; the conv dgrad is the forward tcgen05 conv with role-swapped
// weights, packing.py.)
//
// All reductions are two-stage and deterministic: per-CTA partials in double precision to a workspace, combined in a fixed
// order by a finalize kernel.
#include <algorithm>
#include "common.cuh"

namespace dcs {

constexpr int kTrThreads = 256;

// ------------------------------------------------------------------------------------------------ train-mode complex BN
// pass 1: per-channel sums over pixels of (re, im, re^2, im^2, re*im) -> partial[chunk][c][5] (double)
template <typename T>
__global__ void __launch_bounds__(kTrThreads) cbn_moments_kernel(const T* __restrict__ x, double* __restrict__ partial, int64_t n_pix,
                                                                 int C, int64_t pix_per_cta) {
  __shared__ double red[kTrThreads][5];
  const int lanes = kTrThreads / C, c = threadIdx.x % C, pl = threadIdx.x / C;
  const int64_t p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, n_pix);
  double s[5] = {0, 0, 0, 0, 0};
  if (pl < lanes)
    for (int64_t p = p0 + pl; p < p1; p += lanes) {
      const float2 v = Elem<T>::ldc(x, p * C + c);
      s[0] += v.x; s[1] += v.y; s[2] += (double)v.x * v.x; s[3] += (double)v.y * v.y; s[4] += (double)v.x * v.y;
    }
#pragma unroll
  for (int k = 0; k < 5; ++k) red[threadIdx.x][k] = s[k];
  __syncthreads();
  if (threadIdx.x < C) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double t = 0;
      for (int l = 0; l < lanes; ++l) t += red[l * C + threadIdx.x][k];
      partial[((int64_t)blockIdx.x * C + threadIdx.x) * 5 + k] = t;
    }
  }
}

// finalize: batch statistics -> whitening matrix -> folded affine (the dcs_cbn_apply operand), running-stat update
// (momentum m, unbiased covariance, the eps-inclusive Crr / Cii are what is accumulated), saved statistics for the backward
__global__ void cbn_train_finalize_kernel(const double* __restrict__ partial, int n_chunks, int C, double n, float eps, float momentum,
                                          const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ affine,
                                          float* __restrict__ running_mean, float* __restrict__ running_covar,
                                          long long* __restrict__ num_batches_tracked, float* __restrict__ saved) {
  // one CTA (64 threads) per channel: the chunk partials are summed in a strided fixed order + a fixed tree, thread 0 finishes
  __shared__ double red[64][5];
  const int c = blockIdx.x;
  if (c == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  double s[5] = {0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < n_chunks; k += 64)
#pragma unroll
    for (int j = 0; j < 5; ++j) s[j] += partial[((int64_t)k * C + c) * 5 + j];
#pragma unroll
  for (int j = 0; j < 5; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  for (int o = 32; o; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int j = 0; j < 5; ++j) red[threadIdx.x][j] += red[threadIdx.x + o][j];
    __syncthreads();
  }
  if (threadIdx.x) return;
#pragma unroll
  for (int j = 0; j < 5; ++j) s[j] = red[0][j];
  const double mr = s[0] / n, mi = s[1] / n;
  const double Crr = s[2] / n - mr * mr + (double)eps, Cii = s[3] / n - mi * mi + (double)eps, Cri = s[4] / n - mr * mi;
  const double sd = sqrt(Crr * Cii - Cri * Cri), t = sqrt(Crr + Cii + 2 * sd), ist = 1.0 / (sd * t);
  const double Rrr = (Cii + sd) * ist, Rii = (Crr + sd) * ist, Rri = -Cri * ist;
  const double w0 = weight[3 * c], w1 = weight[3 * c + 1], w2 = weight[3 * c + 2];
  const double A00 = w0 * Rrr + w2 * Rri, A01 = w0 * Rri + w2 * Rii, A10 = w2 * Rrr + w1 * Rri, A11 = w2 * Rri + w1 * Rii;
  float* a = affine + 6 * c;
  a[0] = (float)A00; a[1] = (float)A01; a[2] = (float)A10; a[3] = (float)A11;
  a[4] = (float)(bias[2 * c] - (A00 * mr + A01 * mi));
  a[5] = (float)(bias[2 * c + 1] - (A10 * mr + A11 * mi));
  if (running_mean) {
    const double m = momentum, ub = n / (n - 1.0);
    running_mean[2 * c] = (float)(m * mr + (1 - m) * running_mean[2 * c]);
    running_mean[2 * c + 1] = (float)(m * mi + (1 - m) * running_mean[2 * c + 1]);
    running_covar[3 * c] = (float)(m * Crr * ub + (1 - m) * running_covar[3 * c]);
    running_covar[3 * c + 1] = (float)(m * Cii * ub + (1 - m) * running_covar[3 * c + 1]);
    running_covar[3 * c + 2] = (float)(m * Cri * ub + (1 - m) * running_covar[3 * c + 2]);
  }
  if (saved) {
    float* sv = saved + 8 * c;
    sv[0] = (float)mr; sv[1] = (float)mi; sv[2] = (float)Rrr; sv[3] = (float)Rii; sv[4] = (float)Rri;
    sv[5] = (float)Crr; sv[6] = (float)Cii; sv[7] = (float)Cri;
  }
}

// backward pass 1: per-channel sums over pixels of the eight products of cbn_train_backward:
//   (gr zr, gi zi, gr zi + gi zr, gr, gi, dzr xr, dzi xi, dzr xi + dzi xr)  with xc = x - mean, z = R xc, dz = Wm g
__global__ void __launch_bounds__(kTrThreads) cbn_bwd_sums_kernel(const void* __restrict__ x, int xdt, const float* __restrict__ dy,
                                                                  const float* __restrict__ saved, const float* __restrict__ weight,
                                                                  double* __restrict__ partial, int64_t n_pix, int C, int64_t pix_per_cta) {
  __shared__ double red[kTrThreads][8];
  const int lanes = kTrThreads / C, c = threadIdx.x % C, pl = threadIdx.x / C;
  const int64_t p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, n_pix);
  const float* sv = saved + 8 * c;
  const float mr = sv[0], mi = sv[1], Rrr = sv[2], Rii = sv[3], Rri = sv[4];
  const float w0 = weight[3 * c], w1 = weight[3 * c + 1], w2 = weight[3 * c + 2];
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (pl < lanes)
    for (int64_t p = p0 + pl; p < p1; p += lanes) {
      const float2 xv = ld_c(x, p * C + c, xdt), g = reinterpret_cast<const float2*>(dy)[p * C + c];
      const float xr = xv.x - mr, xi = xv.y - mi;
      const float zr = Rrr * xr + Rri * xi, zi = Rii * xi + Rri * xr;
      const float dzr = w0 * g.x + w2 * g.y, dzi = w2 * g.x + w1 * g.y;
      s[0] += (double)g.x * zr; s[1] += (double)g.y * zi; s[2] += (double)g.x * zi + (double)g.y * zr;
      s[3] += g.x; s[4] += g.y;
      s[5] += (double)dzr * xr; s[6] += (double)dzi * xi; s[7] += (double)dzr * xi + (double)dzi * xr;
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = s[k];
  __syncthreads();
  if (threadIdx.x < C) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double t = 0;
      for (int l = 0; l < lanes; ++l) t += red[l * C + threadIdx.x][k];
      partial[((int64_t)blockIdx.x * C + threadIdx.x) * 8 + k] = t;
    }
  }
}

// backward finalize: dweight (C,3), dbias (C,2) and the per-channel coefficients of pass 2:
//   [dx_re; dx_im] = P [g_re; g_im] + Q [x_re; x_im] + k      (coef[c] = P00 P01 P10 P11 Q00 Q01 Q10 Q11 k0 k1)
// with P = R Wm, Q = (1/n) [[2 dA, dC], [dC, 2 dB]] (the 3x3 Jacobian of the inverse-square-root whitening matrix turns
// (dRrr, dRii, dRri) into (dA, dB, dC)), k = -Q mean - P mean(g)  (the mean of dx is removed).
__global__ void cbn_bwd_finalize_kernel(const double* __restrict__ partial, int n_chunks, int C, double n, const float* __restrict__ saved,
                                        const float* __restrict__ weight, float* __restrict__ dweight, float* __restrict__ dbias,
                                        float* __restrict__ coef) {
  __shared__ double red[64][8];
  const int c = blockIdx.x;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < n_chunks; k += 64)
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += partial[((int64_t)k * C + c) * 8 + j];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  for (int o = 32; o; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[threadIdx.x + o][j];
    __syncthreads();
  }
  if (threadIdx.x) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = red[0][j];
  const float* sv = saved + 8 * c;
  const double mr = sv[0], mi = sv[1], Rrr = sv[2], Rii = sv[3], Rri = sv[4], A = sv[5], B = sv[6], Cc = sv[7];
  const double w0 = weight[3 * c], w1 = weight[3 * c + 1], w2 = weight[3 * c + 2];
  dweight[3 * c] = (float)s[0]; dweight[3 * c + 1] = (float)s[1]; dweight[3 * c + 2] = (float)s[2];
  dbias[2 * c] = (float)s[3]; dbias[2 * c + 1] = (float)s[4];
  const double dRrr = s[5], dRii = s[6], dRri = s[7];
  const double sd = sqrt(A * B - Cc * Cc), t = sqrt(A + B + 2 * sd), u = 1.0 / (sd * t);
  const double s_a = B / (2 * sd), s_b = A / (2 * sd), s_c = -Cc / sd;
  const double t_a = (1 + 2 * s_a) / (2 * t), t_b = (1 + 2 * s_b) / (2 * t), t_c = s_c / t;
  const double u_a = -u * (s_a / sd + t_a / t), u_b = -u * (s_b / sd + t_b / t), u_c = -u * (s_c / sd + t_c / t);
  const double dA = dRrr * (s_a * u + (B + sd) * u_a) + dRii * ((1 + s_a) * u + (A + sd) * u_a) + dRri * (-Cc * u_a);
  const double dB = dRrr * ((1 + s_b) * u + (B + sd) * u_b) + dRii * (s_b * u + (A + sd) * u_b) + dRri * (-Cc * u_b);
  const double dC = dRrr * (s_c * u + (B + sd) * u_c) + dRii * (s_c * u + (A + sd) * u_c) + dRri * (-u - Cc * u_c);
  const double P00 = Rrr * w0 + Rri * w2, P01 = Rrr * w2 + Rri * w1, P10 = Rri * w0 + Rii * w2, P11 = Rri * w2 + Rii * w1;
  const double Q00 = 2 * dA / n, Q01 = dC / n, Q11 = 2 * dB / n;
  const double gmr = s[3] / n, gmi = s[4] / n;
  float* k = coef + 10 * c;
  k[0] = (float)P00; k[1] = (float)P01; k[2] = (float)P10; k[3] = (float)P11;
  k[4] = (float)Q00; k[5] = (float)Q01; k[6] = (float)Q01; k[7] = (float)Q11;
  k[8] = (float)(-(Q00 * mr + Q01 * mi) - (P00 * gmr + P01 * gmi));
  k[9] = (float)(-(Q01 * mr + Q11 * mi) - (P10 * gmr + P11 * gmi));
}

// dx = P dy + Q x + k; with `colsum` the per-channel sums of dx (the bias gradient of the convolution in front: exactly zero in exact
// arithmetic, round-off in fp32 — reported as computed, not assumed) go to colsum[cta][C][2] (a thread keeps ONE channel:
// gridDim.x * 256 is a multiple of C)
__global__ void __launch_bounds__(256) cbn_bwd_apply_kernel(const void* __restrict__ x, int xdt, const float* __restrict__ dy, const float* __restrict__ coef,
                                                            float* __restrict__ dx, int64_t n, int C, double* __restrict__ colsum) {
  __shared__ double red[256][2];
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* k = coef + 10 * (int)(i % C);
    const float2 xv = ld_c(x, i, xdt), g = reinterpret_cast<const float2*>(dy)[i];
    const float2 o = make_float2(k[0] * g.x + k[1] * g.y + k[4] * xv.x + k[5] * xv.y + k[8],
                                 k[2] * g.x + k[3] * g.y + k[6] * xv.x + k[7] * xv.y + k[9]);
    reinterpret_cast<float2*>(dx)[i] = o;
    s0 += o.x; s1 += o.y;
  }
  if (!colsum) return;
  red[threadIdx.x][0] = s0; red[threadIdx.x][1] = s1;
  __syncthreads();
  if (threadIdx.x < C) {
    double t0 = 0.0, t1 = 0.0;
    for (int l = threadIdx.x; l < 256; l += C) { t0 += red[l][0]; t1 += red[l][1]; }
    colsum[((int64_t)blockIdx.x * C + threadIdx.x) * 2] = t0;
    colsum[((int64_t)blockIdx.x * C + threadIdx.x) * 2 + 1] = t1;
  }
}
// conv_r.bias.grad = S.re + S.im, conv_i.bias.grad = S.im - S.re of the convolution in front (complexPyTorch's bias rule), one warp per channel
__global__ void cbn_bwd_colsum_finalize_kernel(const double* __restrict__ colsum, int n_ctas, int C, float* __restrict__ db_r, float* __restrict__ db_i) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double sr = 0.0, si = 0.0;
  for (int k = lane; k < n_ctas; k += 32) { sr += colsum[((int64_t)k * C + c) * 2]; si += colsum[((int64_t)k * C + c) * 2 + 1]; }
#pragma unroll
  for (int o = 16; o; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane == 0) { db_r[c] = (float)(sr + si); db_i[c] = (float)(si - sr); }
}

// ------------------------------------------------------------------------------------------------ SI-SNR value + gradient
// One CTA per batch row (network_functions.py:30-42): s_t = <e, c> c / (<c, c> + eps); ratio = |s_t|^2 / (|e - s_t|^2 + eps) + eps;
// value[row] = 10 log10(ratio); grad = scale / rows * d value / d estimate (oracle/train_oracle.si_snr_backward).
__global__ void __launch_bounds__(512) si_snr_kernel(const float* __restrict__ clean, const float* __restrict__ est, int L, int rows,
                                                     float eps, float scale, float* __restrict__ value, float* __restrict__ grad) {
  __shared__ double red[3][16];
  __shared__ double tot[3];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* c = clean + (int64_t)row * L;
  const float* e = est + (int64_t)row * L;
  double s[3] = {0, 0, 0};   // <e, c>, <c, c>, <e, e>
  for (int i = tid; i < L; i += 512) { const double cv = c[i], ev = e[i]; s[0] += ev * cv; s[1] += cv * cv; s[2] += ev * ev; }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) red[k][warp] = s[k];
  }
  __syncthreads();
  if (tid < 3) { double t = 0; for (int w = 0; w < 16; ++w) t += red[tid][w]; tot[tid] = t; }
  __syncthreads();
  const double dot = tot[0], cc = tot[1], ee = tot[2], ep = eps;
  const double k = dot / (cc + ep);
  const double Tn = k * k * cc, Nn = ee - 2 * k * dot + k * k * cc;     // |s_t|^2, |e - s_t|^2
  const double ratio = Tn / (Nn + ep) + ep;
  if (tid == 0 && value) value[row] = (float)(10.0 * log10(ratio));
  if (!grad) return;
  // d ratio / d e = dTn / (Nn + eps) - Tn dNn / (Nn + eps)^2, dTn = 2 k cc / (cc + eps) c, dNn = 2 (err - <err, c> / (cc + eps) c)
  const double ec = dot - k * cc;                                       // <e - s_t, c>
  const double g0 = (10.0 / log(10.0)) / rows / ratio * scale;
  const double a_c = g0 * (2 * k * cc / (cc + ep) / (Nn + ep) + Tn / ((Nn + ep) * (Nn + ep)) * 2 * (ec / (cc + ep) + k));   // coefficient of c
  const double a_e = -g0 * Tn / ((Nn + ep) * (Nn + ep)) * 2;                                                                 // coefficient of e
  float* g = grad + (int64_t)row * L;
  for (int i = tid; i < L; i += 512) g[i] = (float)(a_c * c[i] + a_e * e[i]);
}

// ------------------------------------------------------------------------------------------------ mask tail adjoint
struct C2 { float x, y; };
__device__ __forceinline__ float2 polar_bwd(float2 s, float2 g, float eps) {   // oracle/train_oracle.polar_roundtrip_backward
  const float rho = sqrtf(s.x * s.x + s.y * s.y), a = s.x + eps;
  const float q = a * a + s.y * s.y;
  if (rho == 0.f || q == 0.f) return g;
  const float ih = rsqrtf(q), c = a * ih, sn = s.y * ih;
  const float drho = g.x * c + g.y * sn, dphi = rho * (-g.x * sn + g.y * c);
  return make_float2(drho * s.x / rho + dphi * (-s.y / q), drho * s.y / rho + dphi * (a / q));
}
__device__ __forceinline__ float2 bound_crm_bwd(float2 m, float2 g, float eps) {   // oracle/train_oracle.bound_crm_backward
  const float rho = sqrtf(m.x * m.x + m.y * m.y);
  const float t = tanhf(rho), a = m.x + eps;
  const float q1 = a * a + m.y * m.y;
  if (rho == 0.f || q1 == 0.f) return make_float2(0.f, 0.f);
  const float ih1 = rsqrtf(q1), c1 = a * ih1, s1 = m.y * ih1;
  const float r1 = t * c1, i1 = t * s1, a2 = r1 + eps;
  const float q2 = a2 * a2 + i1 * i1;
  const float ih2 = q2 > 0.f ? rsqrtf(q2) : 0.f, c2 = q2 > 0.f ? a2 * ih2 : 1.f, s2 = i1 * ih2;
  float dt = g.x * c2 + g.y * s2;
  const float dth2 = t * (-g.x * s2 + g.y * c2);
  const float dr1 = q2 > 0.f ? dth2 * (-i1 / q2) : 0.f, di1 = q2 > 0.f ? dth2 * (a2 / q2) : 0.f;
  dt += dr1 * c1 + di1 * s1;
  const float dth1 = t * (-dr1 * s1 + di1 * c1);
  const float sech2 = 1.f - t * t;
  return make_float2(dt * sech2 * m.x / rho + dth1 * (-m.y / q1), dt * sech2 * m.y / rho + dth1 * (a / q1));
}
// raw: decoder[6] output (un-bounded), Y: noisy spectrogram, gS / gN: iSTFT-adjoint gradients of the clean / noise waveforms
// (gN == nullptr: dc, S = Y M).  d_raw = bound^T bound^T (conj(Y) (polar^T_N gN - polar^T_S gS))   [dcs]
__global__ void mask_tail_bwd_kernel(const float2* __restrict__ raw, const float2* __restrict__ Y, const float2* __restrict__ gS,
                                     const float2* __restrict__ gN, float2* __restrict__ d_raw, int64_t n, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 r = raw[i], y = Y[i];
    const float2 m1 = bound_crm_dev(r, eps, true), m2 = bound_crm_dev(m1, eps, true);
    const float2 prod = cmul(y, m2);
    float2 dprod;
    if (gN) {
      const float2 cl = make_float2(y.x - prod.x, y.y - prod.y);
      const float2 pn = polar_bwd(prod, gN[i], eps), ps = polar_bwd(cl, gS[i], eps);
      dprod = make_float2(pn.x - ps.x, pn.y - ps.y);
    } else {
      dprod = polar_bwd(prod, gS[i], eps);
    }
    // d m2 = conj(Y) * dprod in the (dL/dRe + j dL/dIm) convention
    const float2 dm2 = make_float2(y.x * dprod.x + y.y * dprod.y, y.x * dprod.y - y.y * dprod.x);
    d_raw[i] = bound_crm_bwd(r, bound_crm_bwd(m1, dm2, eps), eps);
  }
}

// ------------------------------------------------------------------------------------------------ decoder input adjoint
// adjoint of cat((d, skip), 1) followed by nearest up-sampling (c_network.py:214-215): g (B, H*uh, W*uw, c0 + c1) complex,
// the dgrad of the up-sampled concatenated tensor -> gd (B, H, W, c0), gskip (B, H, W, c1): sum over each uh x uw block, split.
__global__ void upcat_adjoint_kernel(const float2* __restrict__ g, float2* __restrict__ gd, float2* __restrict__ gskip, int B, int H, int W,
                                     int c0, int c1, int uh, int uw) {
  const int C = c0 + c1;
  const int64_t n = (int64_t)B * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float2 s = make_float2(0.f, 0.f);
    for (int dy = 0; dy < uh; ++dy)
      for (int dx = 0; dx < uw; ++dx) {
        const float2 v = g[(((int64_t)b * H * uh + y * uh + dy) * (W * uw) + x * uw + dx) * C + c];
        s.x += v.x; s.y += v.y;
      }
    const int64_t pix = ((int64_t)b * H + y) * W + x;
    if (c < c0) gd[pix * c0 + c] = s; else gskip[pix * c1 + (c - c0)] = s;
  }
}

static int chunking(int64_t n_pix, int C, int* n_chunks, int64_t* ppc) {
  const int lanes = kTrThreads / C;
  int64_t ctas = std::min<int64_t>((n_pix + lanes * 16 - 1) / (lanes * 16), 4 * (int64_t)num_sms());
  ctas = std::max<int64_t>(ctas, 1);
  *ppc = (n_pix + ctas - 1) / ctas;
  *n_chunks = (int)((n_pix + *ppc - 1) / *ppc);
  return 0;
}

}  // namespace dcs

using namespace dcs;

static bool tr_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" int64_t dcs_cbn_train_workspace_bytes(int64_t n_pix, int channels) {
  if (n_pix <= 0 || !tr_pow2(channels) || channels > 256) return -1;
  int nc; int64_t ppc;
  chunking(n_pix, channels, &nc, &ppc);
  return (int64_t)nc * channels * 8 * sizeof(double) + (int64_t)channels * 16 * sizeof(float);
}

extern "C" int dcs_cbn_train_fwd(const dcs_cbn_train_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->weight && p->bias && p->affine && p->workspace, "dcs_cbn_train_fwd: null pointer");
  DCS_REQUIRE(p->n_pix > 1 && tr_pow2(p->channels) && p->channels <= 256, "dcs_cbn_train_fwd: channels must be a power of two <= 256, n_pix > 1");
  DCS_REQUIRE((p->running_mean == nullptr) == (p->running_covar == nullptr), "dcs_cbn_train_fwd: running_mean / running_covar go together");
  DCS_REQUIRE(is_dtype(p->in_dtype) && is_dtype(p->out_dtype), "dcs_cbn_train_fwd: bad dtype");
  DCS_REQUIRE(p->workspace_bytes >= dcs_cbn_train_workspace_bytes(p->n_pix, p->channels), "dcs_cbn_train_fwd: workspace too small");
  int nc; int64_t ppc;
  chunking(p->n_pix, p->channels, &nc, &ppc);
  double* partial = reinterpret_cast<double*>(p->workspace);
  cudaStream_t s = (cudaStream_t)stream;
  if (p->in_dtype == DCS_F32) cbn_moments_kernel<float><<<nc, kTrThreads, 0, s>>>((const float*)p->x, partial, p->n_pix, p->channels, ppc);
  else if (p->in_dtype == DCS_F16) cbn_moments_kernel<__half><<<nc, kTrThreads, 0, s>>>((const __half*)p->x, partial, p->n_pix, p->channels, ppc);
  else cbn_moments_kernel<__nv_bfloat16><<<nc, kTrThreads, 0, s>>>((const __nv_bfloat16*)p->x, partial, p->n_pix, p->channels, ppc);
  DCS_LAUNCHED();
  cbn_train_finalize_kernel<<<p->channels, 64, 0, s>>>(partial, nc, p->channels, (double)p->n_pix, p->eps, p->momentum, p->weight, p->bias,
                                                                    p->affine, p->running_mean, p->running_covar,
                                                                    reinterpret_cast<long long*>(p->num_batches_tracked), p->saved);
  DCS_LAUNCHED();
  dcs_cbn_params a;
  a.x = p->x; a.y = p->y; a.affine = p->affine; a.n_pix = p->n_pix; a.channels = p->channels; a.act = p->act;
  a.in_dtype = p->in_dtype; a.out_dtype = p->out_dtype;
  return dcs_cbn_apply(&a, stream);
}

extern "C" int dcs_cbn_train_bwd(const dcs_cbn_train_bwd_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->dy && p->dx && p->saved && p->weight && p->dweight && p->dbias && p->workspace, "dcs_cbn_train_bwd: null pointer");
  DCS_REQUIRE(p->n_pix > 1 && tr_pow2(p->channels) && p->channels <= 256, "dcs_cbn_train_bwd: channels must be a power of two <= 256");
  DCS_REQUIRE(is_dtype(p->x_dtype), "dcs_cbn_train_bwd: bad x_dtype");
  DCS_REQUIRE(p->workspace_bytes >= dcs_cbn_train_workspace_bytes(p->n_pix, p->channels), "dcs_cbn_train_bwd: workspace too small");
  int nc; int64_t ppc;
  chunking(p->n_pix, p->channels, &nc, &ppc);
  double* partial = reinterpret_cast<double*>(p->workspace);
  float* coef = reinterpret_cast<float*>(partial + (int64_t)nc * p->channels * 8);
  cudaStream_t s = (cudaStream_t)stream;
  cbn_bwd_sums_kernel<<<nc, kTrThreads, 0, s>>>(p->x, p->x_dtype, p->dy, p->saved, p->weight, partial, p->n_pix, p->channels, ppc);
  DCS_LAUNCHED();
  cbn_bwd_finalize_kernel<<<p->channels, 64, 0, s>>>(partial, nc, p->channels, (double)p->n_pix, p->saved, p->weight, p->dweight, p->dbias, coef);
  DCS_LAUNCHED();
  const int64_t n = p->n_pix * p->channels;
  const bool cs = p->conv_bias_grad_r && p->conv_bias_grad_i;
  // with the column sums the per-CTA partials reuse the (consumed) moment partials: at most nc CTAs
  const int g = (int)std::min<int64_t>((n + 255) / 256, cs ? std::max<int64_t>(1, (int64_t)nc * 4) : (int64_t)num_sms() * 16);
  cbn_bwd_apply_kernel<<<g, 256, 0, s>>>(p->x, p->x_dtype, p->dy, coef, p->dx, n, p->channels, cs ? partial : nullptr);
  DCS_LAUNCHED();
  if (cs) {
    cbn_bwd_colsum_finalize_kernel<<<(p->channels + 7) / 8, 256, 0, s>>>(partial, g, p->channels, p->conv_bias_grad_r, p->conv_bias_grad_i);
    DCS_LAUNCHED();
  }
  return 0;
}

extern "C" int dcs_si_snr(const float* clean, const float* estimate, int rows, int length, float eps, float grad_scale, float* value,
                          float* grad, void* stream) {
  DCS_REQUIRE(clean && estimate && rows > 0 && length > 0 && (value || grad), "dcs_si_snr: bad arguments");
  si_snr_kernel<<<rows, 512, 0, (cudaStream_t)stream>>>(clean, estimate, length, rows, eps, grad_scale, value, grad);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_mask_tail_bwd(const float* net_raw, const float* noisy_spec, const float* g_clean, const float* g_noise, float* d_raw,
                                 int64_t n, float atan2_eps, void* stream) {
  DCS_REQUIRE(net_raw && noisy_spec && g_clean && d_raw && n > 0, "dcs_mask_tail_bwd: bad arguments");
  const int g = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  mask_tail_bwd_kernel<<<g, 256, 0, (cudaStream_t)stream>>>((const float2*)net_raw, (const float2*)noisy_spec, (const float2*)g_clean,
                                                             (const float2*)g_noise, (float2*)d_raw, n, atan2_eps);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_upcat_adjoint(const float* g, float* gd, float* gskip, int batch, int h, int w, int c0, int c1, int up_h, int up_w,
                                 void* stream) {
  DCS_REQUIRE(g && gd && (gskip || c1 == 0) && batch > 0 && h > 0 && w > 0 && c0 > 0 && c1 >= 0 && up_h >= 1 && up_w >= 1,
              "dcs_upcat_adjoint: bad arguments");
  const int64_t n = (int64_t)batch * h * w * (c0 + c1);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  upcat_adjoint_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float2*)g, (float2*)gd, (float2*)gskip, batch, h, w, c0, c1, up_h, up_w);
  DCS_LAUNCHED();
  return 0;
}
