// train_bwd.cu — second slice of the TRAINING step (BASELINE configs[4]): the backward kernels that the
// first slice (train.cu, wgrad_tc.cu, the ADJ mode of stft.cu) left open, all fp32, all reductions two-stage and
// deterministic (per-CTA partials to a workspace, combined in a fixed order).  Contracts: oracle/train_oracle.py.
//   generic weight gradient (implicit GEMM, K = pixels) ... cconv2d_backward / decoder_stage_backward / clinear_backward,
//                                                            and the LSTM's dW_ih / dW_hh (h shifted by one step = a 1-D tap)
//   strided-conv dgrad ..................................... zero insertion (dcs_dilate) + the forward conv kernels with
//                                                            role-swapped weights (train_ops.dgrad_conv)
//   activation / dropout adjoints .......................... c_network.py:113 (ComplexReLU), 148 (ComplexLReLU), 195/203/221
//   attention backward ..................................... attention_backward (c_network.py:53-84, 208-211, 219-220)
//   LSTM forward with saved gates + BPTT ................... lstm_forward_saved / lstm_bptt (c_network.py:12-51)
//   Adam-amsgrad with the global-norm clip ................. c_network.py:229-235, config.py:48-49
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace dcs {

// =============================================================================================== generic wgrad GEMM
// dWp[tap][k][n] = sum over output pixels (b, oh, ow) of x[b, oh*sh + dyo(tap), ow*sw + dxo(tap)][k] * dy[b, oh, ow][n]
// as a tiled fp32 GEMM: M' = ntaps*k2 rows (im2col columns of the forward), N = n2, reduction over pixels, split over CTAs
// (gridDim.z) with partial tiles to a workspace.  256 threads, thread tile RM x RN, CTA tile (16 RM) x (16 RN), 16 pixels
// (one segment of an output row) per stage.
constexpr int kWgKC = 16;

struct WgArgs {
  const float* x; const float* dy; float* ws;
  int batch, in_h, in_w, out_h, out_w, k2, n2, x_pitch, dy_pitch, sh, sw, ntaps, mtot, segs_per_row;
  int64_t n_items, items_per_split;
  int8_t dyo[DCS_MAX_TAPS], dxo[DCS_MAX_TAPS];
};

template <int RM, int RN>
__global__ void __launch_bounds__(256) wgrad_generic_kernel(const WgArgs a) {
  constexpr int TM = 16 * RM, TN = 16 * RN;
  __shared__ __align__(16) float As[kWgKC][TM + 4];
  __shared__ __align__(16) float Bs[kWgKC][TN + 4];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  // A loader: this thread always loads row m' = m0 + tid % TM (fixed tap / channel), pixels kk = tid / TM + (256 / TM) j
  const int am = tid % TM, ak0 = tid / TM;
  const int mrow = m0 + am;
  const bool a_ok = mrow < a.mtot;
  const int tap = a_ok ? mrow / a.k2 : 0, kch = a_ok ? mrow % a.k2 : 0;
  const int dyo = a.dyo[tap], dxo = a.dxo[tap];
  const int bn = tid % TN, bk0 = tid / TN;
  const bool b_ok = n0 + bn < a.n2;
  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
  const int64_t it0 = blockIdx.z * a.items_per_split, it1 = min(it0 + a.items_per_split, a.n_items);
  for (int64_t it = it0; it < it1; ++it) {
    const int seg = (int)(it % a.segs_per_row);
    const int64_t row = it / a.segs_per_row;
    const int oh = (int)(row % a.out_h), b = (int)(row / a.out_h);
    const int ow0 = seg * kWgKC;
    const int ih = oh * a.sh + dyo;
    const bool row_ok = a_ok && (unsigned)ih < (unsigned)a.in_h;
    const float* xrow = a.x + ((int64_t)b * a.in_h + (row_ok ? ih : 0)) * a.in_w * a.x_pitch + kch;
#pragma unroll
    for (int j = 0; j < (kWgKC * TM) / 256; ++j) {
      const int kk = ak0 + (256 / TM) * j;
      const int ow = ow0 + kk, iw = ow * a.sw + dxo;
      float v = 0.f;
      if (row_ok && ow < a.out_w && (unsigned)iw < (unsigned)a.in_w) v = xrow[(int64_t)iw * a.x_pitch];
      As[kk][am] = v;
    }
    const float* dyrow = a.dy + (((int64_t)b * a.out_h + oh) * a.out_w) * a.dy_pitch + n0 + bn;
#pragma unroll
    for (int j = 0; j < (kWgKC * TN) / 256; ++j) {
      const int kk = bk0 + (256 / TN) * j;
      const int ow = ow0 + kk;
      Bs[kk][bn] = (b_ok && ow < a.out_w) ? dyrow[(int64_t)ow * a.dy_pitch] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kWgKC; ++kk) {
      float av[RM], bv[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) av[i] = As[kk][ty * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) bv[j] = Bs[kk][tx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = a.ws + (int64_t)blockIdx.z * a.mtot * a.n2;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= a.mtot) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int n = n0 + tx * RN + j;
      if (n < a.n2) out[(int64_t)m * a.n2 + n] = acc[i][j];
    }
  }
}

// Few-channel layers (n2 <= 32: encoder[0..1], decoder[4..6]; millions of pixels, a few thousand gradient elements): the GEMM
// tiling above wastes its N tile and re-reads x once per tap.  Here a persistent CTA walks 4 x 16-pixel output tiles: the x tile with
// its halo and the dy tile are staged in shared memory once, thread t owns gradient row (tap, k) = blockIdx.y * 256 + t with all
// N2 columns in registers and streams the 64 pixels (one LDS of x + N2 / 4 broadcast LDS.128 of dy per N2 FMAs).  One partial
// [rows][n2] block per CTA, combined by split_reduce_kernel in a fixed order.
constexpr int kWsTH = 4, kWsTW = 16;
struct WsArgs {
  const float* x; const float* dy; float* ws;
  int batch, in_h, in_w, out_h, out_w, k2, n2, x_pitch, dy_pitch, sh, sw, ntaps, rows;
  int min_dy, min_dx, XH, XW, tiles_h, tiles_w, n_tiles;
  int8_t dyo[DCS_MAX_TAPS], dxo[DCS_MAX_TAPS];
};
template <int N2>
__global__ void __launch_bounds__(256) wgrad_small_kernel(const WsArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                    // [XH][XW][k2]
  float* dys = smem + (((size_t)a.XH * a.XW * a.k2 + 3) & ~(size_t)3);      // [TH * TW][N2], 16-byte aligned
  const int tid = threadIdx.x;
  const int row = blockIdx.y * 256 + tid;
  const bool ok = row < a.rows;
  const int tap = ok ? row / a.k2 : 0, kch = ok ? row % a.k2 : 0;
  const int xoff = ((a.dyo[tap] - a.min_dy) * a.XW + (a.dxo[tap] - a.min_dx)) * a.k2 + kch;
  float acc[N2];
#pragma unroll
  for (int j = 0; j < N2; ++j) acc[j] = 0.f;
  const int xt = a.XH * a.XW * a.k2;
  for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int tw = t % a.tiles_w, th = (t / a.tiles_w) % a.tiles_h, b = t / (a.tiles_w * a.tiles_h);
    const int oh0 = th * kWsTH, ow0 = tw * kWsTW;
    const int iy0 = oh0 * a.sh + a.min_dy, ix0 = ow0 * a.sw + a.min_dx;
    __syncthreads();                                   // the previous tile's readers are done
    if ((a.k2 & 3) == 0 && (a.x_pitch & 3) == 0 && ((uintptr_t)a.x & 15) == 0) {
      const int k4 = a.k2 >> 2;
      for (int i = tid; i < xt >> 2; i += 256) {
        const int k = i % k4, px = i / k4;
        const int xx = px % a.XW, yy = px / a.XW;
        const int iy = iy0 + yy, ix = ix0 + xx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)iy < (unsigned)a.in_h && (unsigned)ix < (unsigned)a.in_w)
          v = *reinterpret_cast<const float4*>(a.x + (((int64_t)b * a.in_h + iy) * a.in_w + ix) * a.x_pitch + 4 * k);
        reinterpret_cast<float4*>(xs)[i] = v;
      }
    } else {
      for (int i = tid; i < xt; i += 256) {
        const int k = i % a.k2, px = i / a.k2;
        const int xx = px % a.XW, yy = px / a.XW;
        const int iy = iy0 + yy, ix = ix0 + xx;
        xs[i] = ((unsigned)iy < (unsigned)a.in_h && (unsigned)ix < (unsigned)a.in_w)
                    ? a.x[(((int64_t)b * a.in_h + iy) * a.in_w + ix) * a.x_pitch + k] : 0.f;
      }
    }
    for (int i = tid; i < kWsTH * kWsTW * N2; i += 256) {
      const int n = i % N2, px = i / N2;
      const int oh = oh0 + px / kWsTW, ow = ow0 + px % kWsTW;
      dys[i] = (n < a.n2 && oh < a.out_h && ow < a.out_w) ? a.dy[(((int64_t)b * a.out_h + oh) * a.out_w + ow) * a.dy_pitch + n] : 0.f;
    }
    __syncthreads();
    if (ok) {
#pragma unroll
      for (int r = 0; r < kWsTH; ++r) {
        const float* xr = xs + xoff + (r * a.sh * a.XW) * a.k2;
        const int cstep = a.sw * a.k2;
#pragma unroll 4
        for (int c = 0; c < kWsTW; ++c) {
          const float xv = xr[c * cstep];
          const float* dp = dys + (r * kWsTW + c) * N2;
          if constexpr (N2 >= 4) {
#pragma unroll
            for (int j = 0; j < N2; j += 4) {
              const float4 d4 = *reinterpret_cast<const float4*>(dp + j);
              acc[j] = fmaf(xv, d4.x, acc[j]); acc[j + 1] = fmaf(xv, d4.y, acc[j + 1]);
              acc[j + 2] = fmaf(xv, d4.z, acc[j + 2]); acc[j + 3] = fmaf(xv, d4.w, acc[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < N2; ++j) acc[j] = fmaf(xv, dp[j], acc[j]);
          }
        }
      }
    }
  }
  if (ok) {
    float* out = a.ws + ((int64_t)blockIdx.x * a.rows + row) * a.n2;
#pragma unroll
    for (int j = 0; j < N2; ++j)
      if (j < a.n2) out[j] = acc[j];
  }
}

// dst[i] = sum over splits (fixed order) of ws[z][i]
__global__ void split_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dst, int64_t n, int splits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * n + i];
    dst[i] = s;
  }
}

// fold the four real blocks of dWp[tap][2 cin][2 cout] into conv_r / conv_i .weight.grad (oracle cconv2d_backward):
//   dw_r = dWp[re,re] + dWp[im,im], dw_i = dWp[re -> im] - dWp[im -> re];  ComplexConv2d: (cout, cin, taps);
//   ComplexConvTranspose2d (weights (cin_t, cout_t, k, k), run as the flipped, in/out-swapped conv): the gradient of the
//   equivalent conv weight lands at [ci][co][taps - 1 - t].
__global__ void wgrad_fold_complex_kernel(const float* __restrict__ dwp, int ntaps, int cin, int cout, int transposed,
                                          float* __restrict__ dw_r, float* __restrict__ dw_i) {
  const int n = ntaps * cin * cout, N2 = 2 * cout, K2 = 2 * cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % ntaps, ci = (i / ntaps) % cin, co = i / (ntaps * cin);
    const float* w = dwp + ((int64_t)t * K2 + 2 * ci) * N2 + 2 * co;
    const float rr = w[0], ri = w[1], ir = w[N2], ii = w[N2 + 1];     // [k part][n part]
    const int64_t d = transposed ? ((int64_t)ci * cout + co) * ntaps + (ntaps - 1 - t) : ((int64_t)co * cin + ci) * ntaps + t;
    dw_r[d] = rr + ii;
    dw_i[d] = ri - ir;
  }
}

// dst[c][r] = src[r * pitch + c]  (rows x cols -> cols x rows): real GEMM weight gradients, dWp[k][n] -> W.grad[n][k]
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int pitch) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * pitch + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = t[threadIdx.x][i];
  }
}

// =============================================================================================== plain fp32 GEMM
// C[m][n] = sum_k A[m * lda + k] * B[n * ldb_n + k * ldb_k] (+ bias[n]) (+ C[m][n]): the LSTM projections and their data
// gradients (NT and NN forms through the two B strides).  64 x 64 x 16 tiles, 256 threads, 4 x 4 per thread.
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb_n, int ldb_k,
                                                    const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K,
                                                    int accumulate) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loaders: A tile 64 rows x 16 k: thread -> (row = tid / 4, k = (tid % 4) * 4 .. + 3) (k contiguous in memory)
  const int ar = tid >> 2, ak = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + ak + e, m = m0 + ar;
      As[ak + e][ar] = (m < M && k < K) ? A[(int64_t)m * lda + k] : 0.f;
    }
    if (ldb_k == 1) {           // B rows are k-contiguous (NT): same loader as A
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + ak + e, n = n0 + ar;
        Bs[ak + e][ar] = (n < N && k < K) ? B[(int64_t)n * ldb_n + k] : 0.f;
      }
    } else {                    // n-contiguous (NN): thread -> (k = tid / 16, n = (tid % 16) * 4 .. + 3)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + ty, n = n0 + tx * 4 + e;
        Bs[ty][tx * 4 + e] = (n < N && k < K) ? B[(int64_t)n * ldb_n + (int64_t)k * ldb_k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (accumulate) v += C[(int64_t)m * ldc + n];
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

// =============================================================================================== column sums (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t rows, int cols, int pitch, int64_t rows_per_cta,
                                                     double* __restrict__ partial) {
  // thread -> column c = threadIdx.x % cw (cw = min(cols, 256) rounded), row lanes = 256 / cw; columns beyond 256 loop
  __shared__ double red[256];
  const int64_t r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  for (int cb = 0; cb < cols; cb += 256) {
    const int cw = min(256, cols - cb);
    int lanes = 256 / cw;
    if (lanes < 1) lanes = 1;
    const int c = threadIdx.x % cw, rl = threadIdx.x / cw;
    double s = 0.0;
    if (rl < lanes)
      for (int64_t r = r0 + rl; r < r1; r += lanes) s += x[r * pitch + cb + c];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < cw) {
      double t = 0.0;
      for (int l = 0; l < lanes; ++l) t += red[l * cw + threadIdx.x];
      partial[(int64_t)blockIdx.x * cols + cb + threadIdx.x] = t;
    }
    __syncthreads();
  }
}
// mode 0: out0[c] = S[c] (and out1[c] = S[c] when out1 != NULL);  mode 1 (complex conv bias, columns (co, re/im)):
// out0[co] = S[2co] + S[2co+1] (conv_r.bias.grad), out1[co] = S[2co+1] - S[2co] (conv_i.bias.grad)
__global__ void __launch_bounds__(256) colsum_finalize_kernel(const double* __restrict__ partial, int n_chunks, int cols, int mode, float* __restrict__ out0,
                                                              float* __restrict__ out1) {
  // one warp per output: lanes stride over the chunks (fixed order), xor-tree
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int n_out = mode == 0 ? cols : cols / 2;
  if (i >= n_out) return;
  double sr = 0.0, si = 0.0;
  if (mode == 0) {
    for (int k = lane; k < n_chunks; k += 32) sr += partial[(int64_t)k * cols + i];
  } else {
    for (int k = lane; k < n_chunks; k += 32) { sr += partial[(int64_t)k * cols + 2 * i]; si += partial[(int64_t)k * cols + 2 * i + 1]; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane) return;
  if (mode == 0) {
    out0[i] = (float)sr;
    if (out1) out1[i] = (float)sr;
  } else {
    out0[i] = (float)(sr + si);
    out1[i] = (float)(si - sr);
  }
}

// =============================================================================================== element-wise adjoints
// zero insertion: out (B, in_h, in_w, c) <- dy (B, out_h, out_w, c) at (oh*sh, ow*sw), zero elsewhere (float2 units)
__global__ void dilate_kernel(const float2* __restrict__ dy, float2* __restrict__ out, int B, int out_h, int out_w, int in_h, int in_w,
                              int C, int sh, int sw) {
  const int64_t n = (int64_t)B * in_h * in_w * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % in_w); r /= in_w;
    const int h = (int)(r % in_h);
    const int b = (int)(r / in_h);
    float2 v = make_float2(0.f, 0.f);
    if (h % sh == 0 && w % sw == 0 && h / sh < out_h && w / sw < out_w)
      v = dy[(((int64_t)b * out_h + h / sh) * out_w + w / sw) * C + c];
    out[i] = v;
  }
}

// z (B, h*uh, w*uw, c0 + c1) = nearest up-sampling of cat(d, skip) (the decoder convs' input, materialised for the wgrad)
// two complex channels (one 16-byte load) per thread; c0, c1 even
template <typename TO>
__global__ void upcat_fwd2_kernel(const void* __restrict__ d, const void* __restrict__ skip, int idt, TO* __restrict__ z, int B, int H, int W,
                                  int c0h, int c1h, int uh, int uw) {
  const int Ch = c0h + c1h, HH = H * uh, WW = W * uw;
  const int64_t n = (int64_t)B * HH * WW * Ch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ch);
    int64_t r = i / Ch;
    const int x = (int)(r % WW); r /= WW;
    const int y = (int)(r % HH);
    const int b = (int)(r / HH);
    const int64_t pix = ((int64_t)b * H + y / uh) * W + x / uw;
    const float4 v = c < c0h ? ld_c2(d, pix * c0h + c, idt) : ld_c2(skip, pix * c1h + (c - c0h), idt);
    Elem<TO>::stc(z, 2 * i, make_float2(v.x, v.y));
    Elem<TO>::stc(z, 2 * i + 1, make_float2(v.z, v.w));
  }
}

template <typename TO>
__global__ void upcat_fwd_kernel(const void* __restrict__ d, const void* __restrict__ skip, int idt, TO* __restrict__ z, int B, int H, int W,
                                 int c0, int c1, int uh, int uw) {
  const int C = c0 + c1, HH = H * uh, WW = W * uw;
  const int64_t n = (int64_t)B * HH * WW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int x = (int)(r % WW); r /= WW;
    const int y = (int)(r % HH);
    const int b = (int)(r / HH);
    const int64_t pix = ((int64_t)b * H + y / uh) * W + x / uw;
    Elem<TO>::stc(z, i, c < c0 ? ld_c(d, pix * c0 + c, idt) : ld_c(skip, pix * c1 + (c - c0), idt));
  }
}

// dz = act'(y) * (g0 + g1 + chan_const[b][c]) per real component; y = the activation's OUTPUT (ReLU / LeakyReLU keep the sign)
__global__ void act_bwd_kernel(const void* __restrict__ y, int ydt, const float2* __restrict__ g0, const float2* __restrict__ g1,
                               const float2* __restrict__ cc, float2* __restrict__ dz, int64_t hw_c, int C, int64_t n, int act) {
  const float slope = act == DCS_ACT_RELU ? 0.f : (act == DCS_ACT_LRELU ? 0.01f : 1.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 g = g0[i];
    if (g1) { const float2 t = g1[i]; g.x += t.x; g.y += t.y; }
    if (cc) { const float2 t = cc[(i / hw_c) * C + i % C]; g.x += t.x; g.y += t.y; }
    if (act != DCS_ACT_NONE) {
      const float2 v = ld_c(y, i, ydt);
      g.x *= v.x > 0.f ? 1.f : slope;
      g.y *= v.y > 0.f ? 1.f : slope;
    }
    dz[i] = g;
  }
}

// ---- dropout (torch.nn.Dropout on view_as_real: real and imaginary parts drop independently, c_network.py:195-196).
// Philox4x32-10 counter-based generator: element group i (4 floats) draws from counter (i + offset), key = seed, so the
// backward regenerates the forward's mask from (seed, offset) instead of storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, float scale, uint64_t seed,
                               uint64_t offset) {
  const int64_t n4 = (n + 3) / 4;
  const uint32_t thr = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
  const bool vec = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t c = (uint64_t)i + offset;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    if (vec && 4 * i + 3 < n) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      reinterpret_cast<float4*>(y)[i] = make_float4(rr[0] >= thr ? v.x * scale : 0.f, rr[1] >= thr ? v.y * scale : 0.f,
                                                    rr[2] >= thr ? v.z * scale : 0.f, rr[3] >= thr ? v.w * scale : 0.f);
      continue;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t j = 4 * i + e;
      if (j < n) y[j] = rr[e] >= thr ? x[j] * scale : 0.f;
    }
  }
}

// data gradient of a ComplexConv2d with ONE input channel (encoder[0]: 1 -> 8, k7, stride (2,2)) straight from the raw weights:
// dx(b, ih, iw) = sum over the taps that hit this pixel's stride phase, sum_co [w_r dy.re + w_i dy.im, -w_i dy.re + w_r dy.im].
// (The zero-insertion + forward-kernel route computes 4x the taps and pads N = 2 to 16: 11 ms at batch 32 x 4 s against < 1 ms.)
__global__ void __launch_bounds__(256) cconv_dgrad_cin1_kernel(const float2* __restrict__ dy, const float* __restrict__ w_r, const float* __restrict__ w_i,
                                                               float2* __restrict__ dx, int B, int in_h, int in_w, int out_h, int out_w, int cout,
                                                               int kh, int kw, int sh, int sw) {
  extern __shared__ __align__(16) float2 wsm[];            // [tap][co] = (w_r, w_i)
  const int ntaps = kh * kw;
  for (int i = threadIdx.x; i < ntaps * cout; i += 256) {
    const int t = i / cout, co = i % cout;
    wsm[i] = make_float2(w_r[co * ntaps + t], w_i[co * ntaps + t]);
  }
  __syncthreads();
  const int ph = kh / 2, pw = kw / 2;
  const bool vec = (cout & 1) == 0 && ((uintptr_t)dy & 15) == 0;
  const int64_t n = (int64_t)B * in_h * in_w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int iw = (int)(i % in_w), ih = (int)((i / in_w) % in_h), b = (int)(i / ((int64_t)in_w * in_h));
    float ar = 0.f, ai = 0.f;
    for (int ky = (ih + ph) % sh; ky < kh; ky += sh) {
      const int oh = (ih + ph - ky) / sh;
      if (ih + ph - ky < 0 || oh >= out_h) continue;
      for (int kx = (iw + pw) % sw; kx < kw; kx += sw) {
        const int ow = (iw + pw - kx) / sw;
        if (iw + pw - kx < 0 || ow >= out_w) continue;
        const float2* g = dy + (((int64_t)b * out_h + oh) * out_w + ow) * cout;
        const float2* w = wsm + (ky * kw + kx) * cout;
        if (vec) {                                   // two output channels per 16-byte load (dy pixels and the weight rows are 16-byte aligned)
          const float4* g4 = reinterpret_cast<const float4*>(g);
          const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll 4
          for (int c2 = 0; c2 < cout / 2; ++c2) {
            const float4 gv = g4[c2], wv = w4[c2];
            ar += wv.x * gv.x + wv.y * gv.y + wv.z * gv.z + wv.w * gv.w;
            ai += wv.x * gv.y - wv.y * gv.x + wv.z * gv.w - wv.w * gv.z;
          }
        } else {
          for (int co = 0; co < cout; ++co) {
            const float2 gv = g[co], wv = w[co];
            ar += wv.x * gv.x + wv.y * gv.y;
            ai += wv.x * gv.y - wv.y * gv.x;
          }
        }
      }
    }
    dx[i] = make_float2(ar, ai);
  }
}

template <typename T>
__global__ void dropout_h16_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, float p, float scale, uint64_t seed, uint64_t offset) {
  const int64_t n4 = (n + 3) / 4;
  const uint32_t thr = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t c = (uint64_t)i + offset;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t j = 4 * i + e;
      if (j < n) y[j] = rr[e] >= thr ? from_float<T>(to_float<T>(x[j]) * scale) : from_float<T>(0.f);
    }
  }
}

// =============================================================================================== decoder[6] backward, fused
// decoder[6] = ComplexConvTranspose2d(16 -> 1, k3 s1 p1) on the (2,2) nearest up-sampling of cat(d5, skip6) (c_network.py:214-216).
// With ONE output channel every low-resolution input pixel q sees the same 4 x 4 neighbourhood of the full-resolution gradient dpre:
//   data:   g[q][ci]        = sum_{r,c in -1..2} Weff[ci][r][c] * dpre[2q + (r,c)],  Weff[ci][r][c] = sum_{u + a - 1 = r, v + b - 1 = c} conj(W[ci][a][b])
//   weight: dW[ci][a][b]    = sum_q S_ab[q] * conj(z[q][ci]),  S_ab[q] = sum_{u in {0,1}^2} dpre[2q + u + (a - 1, b - 1)]
//   bias:   db_r = sum (dpre.re + dpre.im), db_i = sum (dpre.im - dpre.re)
// (oracle/train_oracle.decoder_stage_backward for cout = 1, up (2,2)).  Replaces the materialised up-sampled input (2 GB at batch
// 32 x 4 s), the few-channel wgrad, two full-resolution dgrad convolutions and two up-sampling adjoints.  Persistent CTAs over
// 8 x 32-pixel low-resolution tiles; phase 1 thread = pixel (S_ab, g, bias), phase 2 thread = (ci, tap) over the tile's pixels.
constexpr int kD6TH = 8, kD6TW = 32, kD6C = 16;
__global__ void __launch_bounds__(256) dec6_bwd_kernel(const void* __restrict__ d, const void* __restrict__ skip, int xdt, const float2* __restrict__ dpre,
                                                       const float* __restrict__ w_r, const float* __restrict__ w_i, int B, int H, int W, int c0, int c1,
                                                       float2* __restrict__ g_d, float2* __restrict__ g_skip, double* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char d6_smem[];
  float2 (*S)[kD6TH * kD6TW] = reinterpret_cast<float2 (*)[kD6TH * kD6TW]>(d6_smem);                                   // [9][256]
  float2 (*zt)[kD6C + 1] = reinterpret_cast<float2 (*)[kD6C + 1]>(d6_smem + sizeof(float2) * 9 * kD6TH * kD6TW);      // [256][17]
  __shared__ float2 nb[2 * kD6TH + 2][2 * kD6TW + 2];
  __shared__ float2 weff[kD6C][16];
  __shared__ float bred[8][2];
  const int tid = threadIdx.x, Cn = c0 + c1;
  for (int i = tid; i < kD6C * 16; i += 256) {
    const int ci = i / 16, r = (i % 16) / 4 - 1, c = i % 4 - 1;
    float2 acc = make_float2(0.f, 0.f);
    if (ci < Cn)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          const int u = r - a + 1, v = c - b + 1;
          if (u >= 0 && u <= 1 && v >= 0 && v <= 1) { acc.x += w_r[ci * 9 + a * 3 + b]; acc.y -= w_i[ci * 9 + a * 3 + b]; }   // conj(W)
        }
    weff[ci][i % 16] = acc;
  }
  const int tiles_w = (W + kD6TW - 1) / kD6TW, tiles_h = (H + kD6TH - 1) / kD6TH;
  const int n_tiles = B * tiles_h * tiles_w;
  const int HH = 2 * H, WW = 2 * W;
  const int py = tid / kD6TW, px = tid % kD6TW;
  // phase-2 role: output (ci, tap) = tid (tid < 9 * Cn): tap = tid % 9, ci = tid / 9
  const int o_tap = tid % 9, o_ci = tid / 9;
  float2 acc_w = make_float2(0.f, 0.f);
  float acc_b0 = 0.f, acc_b1 = 0.f;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, b = t / (tiles_w * tiles_h);
    const int y0 = th * kD6TH, x0 = tw * kD6TW;
    __syncthreads();
    for (int i = tid; i < (2 * kD6TH + 2) * (2 * kD6TW + 2); i += 256) {
      const int r = i / (2 * kD6TW + 2), c = i % (2 * kD6TW + 2);
      const int yy = 2 * y0 - 1 + r, xx = 2 * x0 - 1 + c;
      nb[r][c] = ((unsigned)yy < (unsigned)HH && (unsigned)xx < (unsigned)WW) ? dpre[((int64_t)b * HH + yy) * WW + xx] : make_float2(0.f, 0.f);
    }
    const int y = y0 + py, x = x0 + px;
    const bool ok = y < H && x < W;
    const int64_t q = ((int64_t)b * H + y) * W + x;
    for (int c = 0; c < Cn; ++c) zt[tid][c] = ok ? (c < c0 ? ld_c(d, q * c0 + c, xdt) : ld_c(skip, q * c1 + (c - c0), xdt)) : make_float2(0.f, 0.f);
    __syncthreads();
    float2 v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) v[r][c] = nb[2 * py + r][2 * px + c];       // dpre[2q + (r - 1, c - 1)]
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)      // S_ab = sum_{u,v in {0,1}} dpre[2q + (u + a - 1, v + bb - 1)] -> v[u + a][v + bb]
        S[a * 3 + bb][tid] = make_float2(v[a][bb].x + v[a][bb + 1].x + v[a + 1][bb].x + v[a + 1][bb + 1].x,
                                         v[a][bb].y + v[a][bb + 1].y + v[a + 1][bb].y + v[a + 1][bb + 1].y);
    if (ok) {
      acc_b0 += v[1][1].x + v[1][2].x + v[2][1].x + v[2][2].x;               // the pixel's own 2 x 2 block: sum of dpre.re
      acc_b1 += v[1][1].y + v[1][2].y + v[2][1].y + v[2][2].y;
      for (int ci = 0; ci < Cn; ++ci) {
        float2 g = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 wv = weff[ci][r * 4 + c], dv = v[r][c];
            g.x += wv.x * dv.x - wv.y * dv.y;
            g.y += wv.x * dv.y + wv.y * dv.x;
          }
        if (ci < c0) g_d[q * c0 + ci] = g; else g_skip[q * c1 + (ci - c0)] = g;
      }
    }
    __syncthreads();
    if (tid < 9 * Cn) {
      float2 a2 = make_float2(0.f, 0.f);
#pragma unroll 8
      for (int p = 0; p < kD6TH * kD6TW; ++p) {
        const float2 sv = S[o_tap][p], zv = zt[p][o_ci];
        a2.x += sv.x * zv.x + sv.y * zv.y;        // S * conj(z)
        a2.y += sv.y * zv.x - sv.x * zv.y;
      }
      acc_w.x += a2.x; acc_w.y += a2.y;
    }
  }
  // per-CTA partials: [9 * Cn complex][bias 2]
  double* out = partial + (int64_t)blockIdx.x * (2 * 9 * kD6C + 2);
  if (tid < 9 * Cn) { out[2 * tid] = acc_w.x; out[2 * tid + 1] = acc_w.y; }
#pragma unroll
  for (int o = 16; o; o >>= 1) { acc_b0 += __shfl_xor_sync(0xffffffffu, acc_b0, o); acc_b1 += __shfl_xor_sync(0xffffffffu, acc_b1, o); }
  if ((tid & 31) == 0) { bred[tid >> 5][0] = acc_b0; bred[tid >> 5][1] = acc_b1; }
  __syncthreads();
  if (tid == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int wv = 0; wv < 8; ++wv) { s0 += bred[wv][0]; s1 += bred[wv][1]; }
    out[2 * 9 * kD6C] = s0; out[2 * 9 * kD6C + 1] = s1;
  }
}
__global__ void dec6_bwd_finalize_kernel(const double* __restrict__ partial, int n_ctas, int Cn, float* __restrict__ dw_r, float* __restrict__ dw_i,
                                         float* __restrict__ db_r, float* __restrict__ db_i) {
  const int i = threadIdx.x;          // (ci, tap) for i < 9 Cn; i == 9 Cn: bias
  if (i > 9 * Cn) return;
  const int stride = 2 * 9 * kD6C + 2;
  const int off = i < 9 * Cn ? 2 * i : 2 * 9 * kD6C;
  double sr = 0.0, si = 0.0;
  for (int c = 0; c < n_ctas; ++c) { sr += partial[(int64_t)c * stride + off]; si += partial[(int64_t)c * stride + off + 1]; }
  if (i < 9 * Cn) {
    const int tap = i % 9, ci = i / 9;
    dw_r[ci * 9 + tap] = (float)sr; dw_i[ci * 9 + tap] = (float)si;
  } else {
    *db_r = (float)(sr + si); *db_i = (float)(si - sr);
  }
}

// =============================================================================================== attention backward
// y = s * u, u = a * x (a: channel gate (B,C), s: spatial gate (B,HW)); oracle/train_oracle.attention_backward.
// G = min(C, 32) lanes cooperate on one pixel, 32 / G pixels per warp.

// pass 1: ds = sum_c conj(u_c) dy_c; dspre = ds (.) s (1 - s) per component
__global__ void __launch_bounds__(256) att_bwd_ds_kernel(const void* __restrict__ x, int xdt, const float2* __restrict__ dy, const float2* __restrict__ gate_c,
                                                         const float2* __restrict__ gate_s, float2* __restrict__ dspre, int hw, int C, int G) {
  __shared__ float2 gs[256];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += 256) gs[c] = gate_c[(int64_t)b * C + c];
  __syncthreads();
  const int sub = threadIdx.x % G, grp = threadIdx.x / G, groups = 256 / G;
  const int64_t base = (int64_t)b * hw;
  for (int pbase = blockIdx.x * groups; pbase < hw; pbase += gridDim.x * groups) {
    const int p = pbase + grp;
    float dr = 0.f, di = 0.f;
    if (p < hw)
      for (int c = sub; c < C; c += G) {
        const float2 u = cmul(gs[c], ld_c(x, (base + p) * C + c, xdt)), g = dy[(base + p) * C + c];
        dr += u.x * g.x + u.y * g.y;          // conj(u) * g
        di += u.x * g.y - u.y * g.x;
      }
    for (int o = G >> 1; o; o >>= 1) { dr += __shfl_xor_sync(0xffffffffu, dr, o); di += __shfl_xor_sync(0xffffffffu, di, o); }
    if (sub == 0 && p < hw) {
      const float2 s = gate_s[base + p];
      dspre[base + p] = make_float2(dr * s.x * (1.f - s.x), di * s.y * (1.f - s.y));
    }
  }
}

// gradient of the 7x7 gate conv's weights: 49 taps x 4 statistics components x 2 gradient components per tile, then folded:
//   dw7_r[ch][tap] = sum_p dsp.re st.re + dsp.im st.im ; dw7_i[ch][tap] = sum_p dsp.im st.re - dsp.re st.im
constexpr int kW7TH = 8, kW7TW = 32;
__global__ void __launch_bounds__(256) att_bwd_w7_kernel(const float4* __restrict__ stats, const float2* __restrict__ dspre, int H, int W,
                                                         int tiles_x, int tiles_y, double* __restrict__ partial) {
  __shared__ float4 st[kW7TH + 6][kW7TW + 6];
  __shared__ float2 dsp[kW7TH][kW7TW];
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
  const int y0 = ty * kW7TH, x0 = tx * kW7TW;
  for (int i = threadIdx.x; i < (kW7TH + 6) * (kW7TW + 6); i += 256) {
    const int r = i / (kW7TW + 6), c = i % (kW7TW + 6);
    const int yy = y0 + r - 3, xx = x0 + c - 3;
    st[r][c] = ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? stats[((int64_t)b * H + yy) * W + xx] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < kW7TH * kW7TW; i += 256) {
    const int r = i / kW7TW, c = i % kW7TW;
    dsp[r][c] = (y0 + r < H && x0 + c < W) ? dspre[((int64_t)b * H + y0 + r) * W + x0 + c] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  // thread -> (tap, 8-column strip of the tile): the four outputs of the tap (mean / max statistic x conv_r / conv_i) from ONE 16-byte
  // statistics load and one 8-byte gradient load per pixel (a thread per output needs a load pair per FMA); the four strips are then
  // combined in shared memory in a fixed order.  Output index = tap + 49 ch + 98 which, as the finalize kernel expects.
  __shared__ float red[4][196];
  if (threadIdx.x < 196) {
    const int tap = threadIdx.x % 49, strip = threadIdx.x / 49;
    const int ky = tap / 7, kx = tap % 7;
    float r_mean = 0.f, r_max = 0.f, i_mean = 0.f, i_max = 0.f;
    for (int r = 0; r < kW7TH; ++r)
#pragma unroll
      for (int c = 0; c < kW7TW / 4; ++c) {
        const int cc = strip * (kW7TW / 4) + c;
        const float4 q = st[r + ky][cc + kx];
        const float2 g = dsp[r][cc];
        r_mean += g.x * q.x + g.y * q.y;  r_max += g.x * q.z + g.y * q.w;
        i_mean += g.y * q.x - g.x * q.y;  i_max += g.y * q.z - g.x * q.w;
      }
    red[strip][tap] = r_mean; red[strip][49 + tap] = r_max; red[strip][98 + tap] = i_mean; red[strip][147 + tap] = i_max;
  }
  __syncthreads();
  if (threadIdx.x < 196)
    partial[(int64_t)tile * 196 + threadIdx.x] = (double)((red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]));
}
// out layout = the reference's (1,2,7,7) tensors: dw7_r[ch*49 + tap], dw7_i[ch*49 + tap].  One CTA per output element: the tiles'
// partials are summed by 256 threads in a strided (fixed) order, then a fixed-shape tree.
__global__ void __launch_bounds__(256) att_bwd_w7_finalize_kernel(const double* __restrict__ partial, int n_tiles, float* __restrict__ dw7_r,
                                                                  float* __restrict__ dw7_i) {
  __shared__ double red[256];
  const int i = blockIdx.x;
  double s = 0.0;
  for (int t = threadIdx.x; t < n_tiles; t += 256) s += partial[(int64_t)t * 196 + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int tap = i % 49, ch = (i / 49) & 1, which = i / 98;
    (which ? dw7_i : dw7_r)[ch * 49 + tap] = (float)red[0];
  }
}

// transposed 7x7 gate conv of the gate gradient: dstats[q] = (d mean.re, d mean.im, d max.re, d max.im) = sum_tap W_tap^H dspre(q - off_tap),
// one thread per pixel on a 16 x 64 tile with a 3-pixel halo of dspre in shared memory
constexpr int kDsTH = 16, kDsTW = 64;
__global__ void __launch_bounds__(256) att_bwd_dstats_kernel(const float2* __restrict__ dspre, const float* __restrict__ w7, float4* __restrict__ dstats,
                                                             int H, int W) {
  __shared__ float2 t[kDsTH + 6][kDsTW + 6];
  __shared__ float4 wq[49];
  const int b = blockIdx.z, y0 = blockIdx.y * kDsTH, x0 = blockIdx.x * kDsTW;
  for (int i = threadIdx.x; i < 49; i += 256) wq[i] = make_float4(w7[i], w7[49 + i], w7[98 + i], w7[147 + i]);   // Wr mean, Wr max, Wi mean, Wi max
  for (int i = threadIdx.x; i < (kDsTH + 6) * (kDsTW + 6); i += 256) {
    const int r = i / (kDsTW + 6), c = i % (kDsTW + 6);
    const int yy = y0 + r - 3, xx = x0 + c - 3;
    t[r][c] = ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? dspre[((int64_t)b * H + yy) * W + xx] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kDsTH * kDsTW; i += 256) {
    const int r = i / kDsTW, c = i % kDsTW;
    if (y0 + r >= H || x0 + c >= W) continue;
    float mr = 0.f, mi = 0.f, xr = 0.f, xi = 0.f;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float2 g = t[r + 6 - ky][c + 6 - kx];          // dspre(q - (ky - 3, kx - 3))
        const float4 wv = wq[ky * 7 + kx];
        mr += wv.x * g.x + wv.z * g.y;  mi += -wv.z * g.x + wv.x * g.y;
        xr += wv.y * g.x + wv.w * g.y;  xi += -wv.w * g.x + wv.y * g.y;
      }
    dstats[((int64_t)b * H + y0 + r) * W + x0 + c] = make_float4(mr, mi, xr, xi);
  }
}

// pass 2: du = conj(s) dy + dmean / C + [argmax] dmax; dx = conj(a) du; da partial sums per CTA
__global__ void __launch_bounds__(256) att_bwd_du_kernel(const void* __restrict__ x, int xdt, const float2* __restrict__ dy, const float2* __restrict__ gate_c,
                                                         const float2* __restrict__ gate_s, const float4* __restrict__ dstats,
                                                         float2* __restrict__ dx, double* __restrict__ da_partial, int H, int W, int C, int G) {
  __shared__ float2 gs[256];
  __shared__ float red[256][2];
  const int b = blockIdx.y, hw = H * W;
  for (int c = threadIdx.x; c < C; c += 256) gs[c] = gate_c[(int64_t)b * C + c];
  __syncthreads();
  const int sub = threadIdx.x % G, grp = threadIdx.x / G, groups = 256 / G;
  const int64_t base = (int64_t)b * hw;
  const int nch = (C + G - 1) / G;      // channels per lane (<= 8 for C <= 256)
  float2 da[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) da[k] = make_float2(0.f, 0.f);
  const float invC = 1.f / (float)C;
  for (int pbase = blockIdx.x * groups; pbase < hw; pbase += gridDim.x * groups) {
    const int p = pbase + grp;
    const bool ok = p < hw;
    // u, and the arg max of Re u / Im u over the channels (first index on ties, like torch.max)
    float2 uv[8], gv[8], xv8[8];
    float bre = -INFINITY, bim = -INFINITY;
    int are = 0x7fffffff, aim = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = sub + k * G;
      if (k < nch && ok && c < C) {
        xv8[k] = ld_c(x, (base + p) * C + c, xdt);
        uv[k] = cmul(gs[c], xv8[k]);
        gv[k] = dy[(base + p) * C + c];
        if (uv[k].x > bre) { bre = uv[k].x; are = c; }
        if (uv[k].y > bim) { bim = uv[k].y; aim = c; }
      }
    }
    for (int o = G >> 1; o; o >>= 1) {
      const float ore = __shfl_xor_sync(0xffffffffu, bre, o), oim = __shfl_xor_sync(0xffffffffu, bim, o);
      const int oare = __shfl_xor_sync(0xffffffffu, are, o), oaim = __shfl_xor_sync(0xffffffffu, aim, o);
      if (ore > bre || (ore == bre && oare < are)) { bre = ore; are = oare; }
      if (oim > bim || (oim == bim && oaim < aim)) { bim = oim; aim = oaim; }
    }
    if (ok) {
      const float2 s = gate_s[base + p];
      const float4 ds = dstats[base + p];
      const float mr = ds.x * invC, mi = ds.y * invC;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = sub + k * G;
        if (k < nch && c < C) {
          const float2 g = gv[k];
          float2 du = make_float2(s.x * g.x + s.y * g.y + mr, s.x * g.y - s.y * g.x + mi);   // conj(s) g + dmean / C
          if (c == are) du.x += ds.z;
          if (c == aim) du.y += ds.w;
          const float2 xv = xv8[k], a = gs[c];
          da[k].x += xv.x * du.x + xv.y * du.y;        // conj(x) du
          da[k].y += xv.x * du.y - xv.y * du.x;
          dx[(base + p) * C + c] = make_float2(a.x * du.x + a.y * du.y, a.x * du.y - a.y * du.x);   // conj(a) du
        }
      }
    }
  }
  // CTA reduction of da over the pixel groups (fixed order), one partial row per CTA
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k >= nch) break;
    __syncthreads();
    red[threadIdx.x][0] = da[k].x; red[threadIdx.x][1] = da[k].y;
    __syncthreads();
    const int c = threadIdx.x + k * G;
    if (threadIdx.x < G && c < C) {
      double sr = 0.0, si = 0.0;
      for (int g = 0; g < groups; ++g) { sr += red[g * G + threadIdx.x][0]; si += red[g * G + threadIdx.x][1]; }
      double* o = da_partial + (((int64_t)b * gridDim.x + blockIdx.x) * C + c) * 2;
      o[0] = sr; o[1] = si;
    }
  }
}

// Few-channel tensors (C = 8, 16: the largest ones): ONE THREAD PER PIXEL with all C channels in registers — 16-byte loads, no
// shuffles, a serial arg max — instead of C lanes per pixel (the C = 8 launches ran at a tenth of the HBM rate: 12 shuffles and a
// strided loop per element).
template <int C>
__global__ void __launch_bounds__(256) att_bwd_ds_small_kernel(const void* __restrict__ x, int xdt, const float4* __restrict__ dy, const float2* __restrict__ gate_c,
                                                               const float2* __restrict__ gate_s, float2* __restrict__ dspre, int hw) {
  __shared__ float2 gs[C];
  const int b = blockIdx.y;
  if (threadIdx.x < C) gs[threadIdx.x] = gate_c[(int64_t)b * C + threadIdx.x];
  __syncthreads();
  const int64_t base = (int64_t)b * hw;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < hw; p += gridDim.x * 256) {
    float dr = 0.f, di = 0.f;
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      const float4 xv = ld_c2(x, (base + p) * (C / 2) + c / 2, xdt), g = dy[(base + p) * (C / 2) + c / 2];
      const float2 u0 = cmul(gs[c], make_float2(xv.x, xv.y)), u1 = cmul(gs[c + 1], make_float2(xv.z, xv.w));
      dr += u0.x * g.x + u0.y * g.y + u1.x * g.z + u1.y * g.w;
      di += u0.x * g.y - u0.y * g.x + u1.x * g.w - u1.y * g.z;
    }
    const float2 sg = gate_s[base + p];
    dspre[base + p] = make_float2(dr * sg.x * (1.f - sg.x), di * sg.y * (1.f - sg.y));
  }
}

template <int C>
__global__ void __launch_bounds__(256) att_bwd_du_small_kernel(const void* __restrict__ x, int xdt, const float4* __restrict__ dy, const float2* __restrict__ gate_c,
                                                               const float2* __restrict__ gate_s, const float4* __restrict__ dstats,
                                                               float4* __restrict__ dx, double* __restrict__ da_partial, int hw) {
  __shared__ float2 gs[C];
  __shared__ float red[8][C][2];
  const int b = blockIdx.y, tid = threadIdx.x;
  if (tid < C) gs[tid] = gate_c[(int64_t)b * C + tid];
  __syncthreads();
  const int64_t base = (int64_t)b * hw;
  const float invC = 1.f / (float)C;
  float2 da[C];
#pragma unroll
  for (int c = 0; c < C; ++c) da[c] = make_float2(0.f, 0.f);
  for (int p = blockIdx.x * 256 + tid; p < hw; p += gridDim.x * 256) {
    float2 xv[C], gv[C];
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      const float4 a4 = ld_c2(x, (base + p) * (C / 2) + c / 2, xdt), g4 = dy[(base + p) * (C / 2) + c / 2];
      xv[c] = make_float2(a4.x, a4.y); xv[c + 1] = make_float2(a4.z, a4.w);
      gv[c] = make_float2(g4.x, g4.y); gv[c + 1] = make_float2(g4.z, g4.w);
    }
    float bre = -INFINITY, bim = -INFINITY;
    int are = 0, aim = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {                        // first index on ties, like torch.max
      const float2 u = cmul(gs[c], xv[c]);
      if (u.x > bre) { bre = u.x; are = c; }
      if (u.y > bim) { bim = u.y; aim = c; }
    }
    const float2 s = gate_s[base + p];
    const float4 ds = dstats[base + p];
    const float mr = ds.x * invC, mi = ds.y * invC;
    float2 o[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float2 g = gv[c], a = gs[c];
      float2 du = make_float2(s.x * g.x + s.y * g.y + mr, s.x * g.y - s.y * g.x + mi);
      if (c == are) du.x += ds.z;
      if (c == aim) du.y += ds.w;
      da[c].x += xv[c].x * du.x + xv[c].y * du.y;
      da[c].y += xv[c].x * du.y - xv[c].y * du.x;
      o[c] = make_float2(a.x * du.x + a.y * du.y, a.x * du.y - a.y * du.x);
    }
#pragma unroll
    for (int c = 0; c < C; c += 2) dx[(base + p) * (C / 2) + c / 2] = make_float4(o[c].x, o[c].y, o[c + 1].x, o[c + 1].y);
  }
  // da: warp xor-tree, then the 8 warps in a fixed order
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float vr = da[c].x, vi = da[c].y;
#pragma unroll
    for (int o2 = 16; o2; o2 >>= 1) { vr += __shfl_xor_sync(0xffffffffu, vr, o2); vi += __shfl_xor_sync(0xffffffffu, vi, o2); }
    if ((tid & 31) == 0) { red[tid >> 5][c][0] = vr; red[tid >> 5][c][1] = vi; }
  }
  __syncthreads();
  if (tid < C) {
    double sr = 0.0, si = 0.0;
    for (int w = 0; w < 8; ++w) { sr += red[w][tid][0]; si += red[w][tid][1]; }
    double* o = da_partial + (((int64_t)b * gridDim.x + blockIdx.x) * C + tid) * 2;
    o[0] = sr; o[1] = si;
  }
}

// the gate MLP's backward, one CTA per image: da -> dfc = 2 da (.) a (1 - a) -> W2^T, crelu mask, W1^T -> davg;
// chan_const[b][c] = davg / (H W); per-image weight gradients to `wpart` (B, 4 R C): [dw1_r (R,C) | dw1_i | dw2_r (C,R) | dw2_i]
__global__ void __launch_bounds__(256) att_bwd_gate_kernel(const double* __restrict__ da_partial, int n_chunks, const float2* __restrict__ gate_c,
                                                           const long long* __restrict__ sums, float inv_hw, int C, int R,
                                                           const float* __restrict__ w1_r, const float* __restrict__ w1_i,
                                                           const float* __restrict__ w2_r, const float* __restrict__ w2_i,
                                                           float2* __restrict__ chan_const, float* __restrict__ wpart) {
  __shared__ float2 avg[256], dfc[256], hid_pre[16], hid[16], dhp[16];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) {
    double sr = 0.0, si = 0.0;
    for (int k = 0; k < n_chunks; ++k) {
      const double* o = da_partial + (((int64_t)b * n_chunks + k) * C + c) * 2;
      sr += o[0]; si += o[1];
    }
    const float2 a = gate_c[(int64_t)b * C + c];
    dfc[c] = make_float2(2.f * (float)sr * a.x * (1.f - a.x), 2.f * (float)si * a.y * (1.f - a.y));
    avg[c] = make_float2(pool_mean(sums, ((int64_t)b * C + c) * 2, inv_hw), pool_mean(sums, ((int64_t)b * C + c) * 2 + 1, inv_hw));
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < R; r += 8) {           // hidden pre-activation (as the forward computes it) and dhid
    float re = 0.f, im = 0.f, dr = 0.f, di = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float wr = w1_r[r * C + c], wi = w1_i[r * C + c];
      re += wr * avg[c].x - wi * avg[c].y;
      im += wr * avg[c].y + wi * avg[c].x;
      const float vr = w2_r[c * R + r], vi = w2_i[c * R + r];
      dr += vr * dfc[c].x + vi * dfc[c].y;
      di += -vi * dfc[c].x + vr * dfc[c].y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o);
      dr += __shfl_xor_sync(0xffffffffu, dr, o); di += __shfl_xor_sync(0xffffffffu, di, o);
    }
    if (lane == 0) {
      hid_pre[r] = make_float2(re, im);
      hid[r] = make_float2(fmaxf(re, 0.f), fmaxf(im, 0.f));
      dhp[r] = make_float2(re > 0.f ? dr : 0.f, im > 0.f ? di : 0.f);
    }
  }
  __syncthreads();
  float* wp = wpart + (int64_t)b * 4 * R * C;
  for (int i = tid; i < R * C; i += 256) {
    {   // dw1 (R, C): dY = dhp[r], X = avg[c]
      const int r = i / C, c = i % C;
      wp[i] = dhp[r].x * avg[c].x + dhp[r].y * avg[c].y;
      wp[R * C + i] = dhp[r].y * avg[c].x - dhp[r].x * avg[c].y;
    }
    {   // dw2 (C, R): dY = dfc[c], X = hid[r]
      const int c = i / R, r = i % R;
      wp[2 * R * C + i] = dfc[c].x * hid[r].x + dfc[c].y * hid[r].y;
      wp[3 * R * C + i] = dfc[c].y * hid[r].x - dfc[c].x * hid[r].y;
    }
  }
  const float k = inv_hw;
  for (int c = tid; c < C; c += 256) {
    float re = 0.f, im = 0.f;
    for (int r = 0; r < R; ++r) {
      const float wr = w1_r[r * C + c], wi = w1_i[r * C + c];
      re += wr * dhp[r].x + wi * dhp[r].y;
      im += -wi * dhp[r].x + wr * dhp[r].y;
    }
    chan_const[(int64_t)b * C + c] = make_float2(re * k, im * k);
  }
}
// dst[i] = sum_b src[b][i] in a fixed order
__global__ void batch_reduce_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += src[(int64_t)b * n + i];
    dst[i] = s;
  }
}

// =============================================================================================== LSTM (training form)
// One CTA per (sequence q, direction d); 4H = 256 threads, thread r owns gate row r (gate r / H of unit r % H) with its W_hh
// row in registers; q uses the weights of group q / (n_seq / n_groups).  pre / gates: (Q, S, 2, 4H); h / cells: (Q, S, 2, H).
template <int H>
__global__ void __launch_bounds__(4 * H) lstm_train_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ w_hh, int n_seq, int n_groups,
                                                               int S, float* __restrict__ h_out, float* __restrict__ gates, float* __restrict__ cells) {
  constexpr int G4 = 4 * H;
  __shared__ float hs[H];
  __shared__ float act[G4];
  const int q = blockIdx.x, d = blockIdx.y, r = threadIdx.x;
  const int grp = q / (n_seq / n_groups);
  float w[H];
  const float* wrow = w_hh + (((int64_t)grp * 2 + d) * G4 + r) * H;
#pragma unroll
  for (int k = 0; k < H; ++k) w[k] = wrow[k];
  if (r < H) hs[r] = 0.f;
  float c = 0.f;
  __syncthreads();
  float a_next = pre[(((int64_t)q * S + (d ? S - 1 : 0)) * 2 + d) * G4 + r];
  for (int step = 0; step < S; ++step) {
    const int t = d ? S - 1 - step : step;
    const int64_t row = ((int64_t)q * S + t) * 2 + d;
    float a = a_next;
    if (step + 1 < S) a_next = pre[(((int64_t)q * S + (d ? t - 1 : t + 1)) * 2 + d) * G4 + r];   // prefetch: off the serial chain
#pragma unroll
    for (int k = 0; k < H; ++k) a = fmaf(w[k], hs[k], a);
    const float v = (r / H == 2) ? tanhf(a) : sigmoidf_(a);
    act[r] = v;
    gates[row * G4 + r] = v;
    __syncthreads();
    if (r < H) {
      c = act[H + r] * c + act[r] * act[2 * H + r];
      const float hv = act[3 * H + r] * tanhf(c);
      hs[r] = hv;
      h_out[row * H + r] = hv;
      cells[row * H + r] = c;
    }
    __syncthreads();
  }
}

// BPTT (oracle/train_oracle.lstm_bptt): reverse of the forward's time order; thread r computes da of its gate row, then the
// CTA forms dh_next = da W_hh (thread (k, part) sums 64 rows of column k held in registers).  dpre (Q, S, 2, 4H).
template <int H>
__global__ void __launch_bounds__(4 * H) lstm_train_bwd_kernel(const float* __restrict__ w_hh, const float* __restrict__ gates,
                                                               const float* __restrict__ cells, const float* __restrict__ dh_out, int n_seq,
                                                               int n_groups, int S, float* __restrict__ dpre) {
  constexpr int G4 = 4 * H;
  __shared__ float da_s[G4];
  __shared__ float part[4][H];
  const int q = blockIdx.x, d = blockIdx.y, r = threadIdx.x;
  const int grp = q / (n_seq / n_groups);
  const int gate = r / H, u = r % H;
  const int kcol = r % H, prt = r / H;     // dh_next role: column kcol, rows prt*H .. prt*H + H - 1
  float wt[H];
  const float* wbase = w_hh + ((int64_t)grp * 2 + d) * G4 * H;
#pragma unroll
  for (int k = 0; k < H; ++k) wt[k] = wbase[(int64_t)(prt * H + k) * H + kcol];
  float dc_next = 0.f, dh_next = 0.f;
  // operands of the step being processed are loaded one step ahead (they do not depend on the recurrence)
  auto row_of = [&](int t) { return ((int64_t)q * S + t) * 2 + d; };
  int t = d ? 0 : S - 1;
  float gi, gf, gg, go, cv, dho;
  {
    const float* g = gates + row_of(t) * G4;
    gi = g[u]; gf = g[H + u]; gg = g[2 * H + u]; go = g[3 * H + u];
    cv = cells[row_of(t) * H + u];
    dho = dh_out[row_of(t) * H + u];
  }
  for (int step = 0; step < S; ++step) {
    const int tp = d ? t + 1 : t - 1;                       // the step processed before t in the forward = the next one here
    const bool more = step + 1 < S;
    float ngi = 0.f, ngf = 0.f, ngg = 0.f, ngo = 0.f, ncv = 0.f, ndho = 0.f;
    if (more) {
      const float* g = gates + row_of(tp) * G4;
      ngi = g[u]; ngf = g[H + u]; ngg = g[2 * H + u]; ngo = g[3 * H + u];
      ncv = cells[row_of(tp) * H + u];
      ndho = dh_out[row_of(tp) * H + u];
    }
    const float c_prev = ncv;                               // zero at the first forward step
    const float tc = tanhf(cv);
    const float dh = dho + dh_next;
    const float dc = dh * go * (1.f - tc * tc) + dc_next;
    float da;
    if (gate == 0) da = dc * gg * gi * (1.f - gi);
    else if (gate == 1) da = dc * c_prev * gf * (1.f - gf);
    else if (gate == 2) da = dc * gi * (1.f - gg * gg);
    else da = dh * tc * go * (1.f - go);
    dc_next = dc * gf;
    dpre[row_of(t) * G4 + r] = da;
    da_s[r] = da;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) s = fmaf(da_s[prt * H + k], wt[k], s);
    part[prt][kcol] = s;
    __syncthreads();
    dh_next = part[0][u] + part[1][u] + part[2][u] + part[3][u];
    gi = ngi; gf = ngf; gg = ngg; go = ngo; cv = ncv; dho = ndho; t = tp;
  }
}

// (rows, D, 2) interleaved complex <-> (2, rows, D) planes
__global__ void cplx_split_kernel(const void* __restrict__ x, int xdt, float* __restrict__ planes, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = ld_c(x, i, xdt);
    planes[i] = v.x; planes[n + i] = v.y;
  }
}
__global__ void cplx_merge_kernel(const float* __restrict__ planes, float2* __restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = make_float2(planes[i], planes[n + i]);
}
// ComplexLSTM combine (c_network.py:39-47): h (2 lstm, 2 part, n): out.re = R(re) - I(im), out.im = R(im) + I(re); and its adjoint
__global__ void clstm_combine_kernel(const float* __restrict__ h, float2* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = make_float2(h[i] - h[3 * n + i], h[n + i] + h[2 * n + i]);
}
__global__ void clstm_combine_bwd_kernel(const float2* __restrict__ dout, float* __restrict__ dh, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 g = dout[i];
    dh[i] = g.x; dh[n + i] = g.y; dh[2 * n + i] = g.y; dh[3 * n + i] = -g.x;
  }
}

// =============================================================================================== optimizer
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += (double)x[i] * x[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += red[w]; partial[blockIdx.x] = t; }
}
__global__ void sumsq_finalize_kernel(const double* __restrict__ partial, int n, double* __restrict__ out, int accumulate) {
  double t = accumulate ? *out : 0.0;
  for (int i = 0; i < n; ++i) t += partial[i];
  *out = t;
}
// torch.optim.Adam(amsgrad=True, weight_decay = L2 added to the gradient) on flat buffers; the gradient is first scaled by
// grad_scale (1 / world size) and by the global-norm clip coefficient min(1, max_norm / (norm + 1e-6)) with norm =
// grad_scale * sqrt(*grad_sumsq) (torch.nn.utils.clip_grad_norm_, what Lightning's gradient_clip_algorithm "norm" calls).
__global__ void adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                    float* __restrict__ vmax, int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                    const double* __restrict__ grad_sumsq, float max_norm, float grad_scale) {
  float coef = grad_scale;
  if (grad_sumsq && max_norm > 0.f) {
    const float norm = grad_scale * (float)sqrt(*grad_sumsq);
    coef *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step = lr / bc1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gr = g[i] * coef + wd * pv;
    const float mv = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    const float vm = fmaxf(vmax[i], vv);
    m[i] = mv; v[i] = vv; vmax[i] = vm;
    p[i] = pv - step * mv / (sqrtf(vm) / bc2_sqrt + eps);
  }
}

// packed operand = signed sum of up to 4 source elements (the linear map raw weights -> kernel operand layouts, built once on
// the host as an index table): dst[i] = sum_j sign_j src[idx_j]; idx < 0 = no term
__global__ void gather_pack_kernel(const float* __restrict__ src, const int4* __restrict__ idx, const char4* __restrict__ sgn, void* __restrict__ dst,
                                   int64_t n, int out_dtype) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 id = idx[i];
    const char4 sg = sgn[i];
    float s = 0.f;
    if (id.x >= 0) s += (float)sg.x * src[id.x];
    if (id.y >= 0) s += (float)sg.y * src[id.y];
    if (id.z >= 0) s += (float)sg.z * src[id.z];
    if (id.w >= 0) s += (float)sg.w * src[id.w];
    if (out_dtype == DCS_F32) reinterpret_cast<float*>(dst)[i] = s;
    else if (out_dtype == 3) {        // fp32 rounded to tf32 (nearest even): the tensor core would otherwise truncate the low 13 bits
      uint32_t u = __float_as_uint(s);
      u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
      reinterpret_cast<float*>(dst)[i] = __uint_as_float(u);
    } else if (out_dtype == DCS_F16) reinterpret_cast<__half*>(dst)[i] = from_float<__half>(s);
    else reinterpret_cast<__nv_bfloat16*>(dst)[i] = from_float<__nv_bfloat16>(s);
  }
}

static int ew_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16)); }

struct WgPlan { int rm, rn, mt, nt, splits; int64_t n_items, items_per_split; int segs; int small, N2, workers, row_groups; size_t smem; WsArgs sa; };
static WgPlan wg_plan(const dcs_wgrad_params* p) {
  WgPlan w;
  memset(&w, 0, sizeof(w));
  // few-channel path
  if (p->n2 <= 32 && p->ntaps * p->k2 >= 64 && !getenv("DCS_WGRAD_NO_SMALL")) {
    WsArgs& a = w.sa;
    int mn_y = 127, mx_y = -127, mn_x = 127, mx_x = -127;
    for (int t = 0; t < p->ntaps; ++t) {
      mn_y = std::min<int>(mn_y, p->dy_off[t]); mx_y = std::max<int>(mx_y, p->dy_off[t]);
      mn_x = std::min<int>(mn_x, p->dx_off[t]); mx_x = std::max<int>(mx_x, p->dx_off[t]);
    }
    a.min_dy = mn_y; a.min_dx = mn_x;
    a.XH = (kWsTH - 1) * p->stride_h + (mx_y - mn_y) + 1;
    a.XW = (kWsTW - 1) * p->stride_w + (mx_x - mn_x) + 1;
    w.N2 = p->n2 <= 2 ? 2 : (p->n2 <= 16 ? 16 : 32);
    w.smem = ((((size_t)a.XH * a.XW * p->k2 + 3) & ~(size_t)3) + (size_t)kWsTH * kWsTW * w.N2) * sizeof(float);
    if (w.smem <= 96 * 1024) {
      w.small = 1;
      a.rows = p->ntaps * p->k2;
      w.row_groups = (a.rows + 255) / 256;
      a.tiles_h = (p->out_h + kWsTH - 1) / kWsTH; a.tiles_w = (p->out_w + kWsTW - 1) / kWsTW;
      a.n_tiles = p->batch * a.tiles_h * a.tiles_w;
      w.workers = std::max(1, std::min(a.n_tiles, (3 * num_sms() + w.row_groups - 1) / w.row_groups));
      w.splits = w.workers;
      return w;
    }
  }
  const int mtot = p->ntaps * p->k2;
  w.rn = p->n2 >= 48 ? 4 : 1;
  w.rm = 4;
  const int TM = 16 * w.rm, TN = 16 * w.rn;
  w.mt = (mtot + TM - 1) / TM;
  w.nt = (p->n2 + TN - 1) / TN;
  w.segs = (p->out_w + kWgKC - 1) / kWgKC;
  w.n_items = (int64_t)p->batch * p->out_h * w.segs;
  const int64_t tiles = (int64_t)w.mt * w.nt;
  int64_t splits = std::max<int64_t>(1, (4LL * num_sms() + tiles - 1) / tiles);
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, w.n_items / 4));
  splits = std::min<int64_t>(splits, 1024);
  w.items_per_split = (w.n_items + splits - 1) / splits;
  w.splits = (int)((w.n_items + w.items_per_split - 1) / w.items_per_split);
  return w;
}

}  // namespace dcs

using namespace dcs;

extern "C" int64_t dcs_wgrad_workspace_bytes(const dcs_wgrad_params* p) {
  if (!p || p->ntaps <= 0 || p->ntaps > DCS_MAX_TAPS || p->k2 <= 0 || p->n2 <= 0 || p->batch <= 0 || p->out_h <= 0 || p->out_w <= 0) return -1;
  const WgPlan w = wg_plan(p);
  return (int64_t)w.splits * p->ntaps * p->k2 * p->n2 * sizeof(float);
}

extern "C" int dcs_wgrad(const dcs_wgrad_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->dy && p->dwp && p->workspace, "dcs_wgrad: null pointer");
  DCS_REQUIRE(p->ntaps > 0 && p->ntaps <= DCS_MAX_TAPS && p->k2 > 0 && p->n2 > 0 && p->batch > 0 && p->in_h > 0 && p->in_w > 0 && p->out_h > 0 &&
              p->out_w > 0 && p->stride_h > 0 && p->stride_w > 0 && p->x_pitch >= p->k2 && p->dy_pitch >= p->n2, "dcs_wgrad: bad shape");
  DCS_REQUIRE(p->workspace_bytes >= dcs_wgrad_workspace_bytes(p), "dcs_wgrad: workspace too small");
  const WgPlan w = wg_plan(p);
  if (w.small) {
    WsArgs sa = w.sa;
    sa.x = p->x; sa.dy = p->dy; sa.ws = reinterpret_cast<float*>(p->workspace);
    sa.batch = p->batch; sa.in_h = p->in_h; sa.in_w = p->in_w; sa.out_h = p->out_h; sa.out_w = p->out_w; sa.k2 = p->k2; sa.n2 = p->n2;
    sa.x_pitch = p->x_pitch; sa.dy_pitch = p->dy_pitch; sa.sh = p->stride_h; sa.sw = p->stride_w; sa.ntaps = p->ntaps;
    for (int t = 0; t < p->ntaps; ++t) { sa.dyo[t] = p->dy_off[t]; sa.dxo[t] = p->dx_off[t]; }
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(w.workers, w.row_groups);
    if (w.N2 == 2) {
      DCS_CUDA(cudaFuncSetAttribute(wgrad_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
      wgrad_small_kernel<2><<<grid, 256, w.smem, s>>>(sa);
    } else if (w.N2 == 16) {
      DCS_CUDA(cudaFuncSetAttribute(wgrad_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
      wgrad_small_kernel<16><<<grid, 256, w.smem, s>>>(sa);
    } else {
      DCS_CUDA(cudaFuncSetAttribute(wgrad_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
      wgrad_small_kernel<32><<<grid, 256, w.smem, s>>>(sa);
    }
    DCS_LAUNCHED();
    const int64_t n = (int64_t)sa.rows * p->n2;
    split_reduce_kernel<<<ew_grid(n), 256, 0, s>>>(sa.ws, p->dwp, n, w.workers);
    DCS_LAUNCHED();
    return 0;
  }
  WgArgs a;
  a.x = p->x; a.dy = p->dy; a.ws = reinterpret_cast<float*>(p->workspace);
  a.batch = p->batch; a.in_h = p->in_h; a.in_w = p->in_w; a.out_h = p->out_h; a.out_w = p->out_w; a.k2 = p->k2; a.n2 = p->n2;
  a.x_pitch = p->x_pitch; a.dy_pitch = p->dy_pitch; a.sh = p->stride_h; a.sw = p->stride_w; a.ntaps = p->ntaps; a.mtot = p->ntaps * p->k2;
  a.segs_per_row = w.segs; a.n_items = w.n_items; a.items_per_split = w.items_per_split;
  for (int t = 0; t < p->ntaps; ++t) { a.dyo[t] = p->dy_off[t]; a.dxo[t] = p->dx_off[t]; }
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(w.mt, w.nt, w.splits);
  if (w.rn == 4) wgrad_generic_kernel<4, 4><<<grid, 256, 0, s>>>(a);
  else wgrad_generic_kernel<4, 1><<<grid, 256, 0, s>>>(a);
  DCS_LAUNCHED();
  const int64_t n = (int64_t)a.mtot * p->n2;
  split_reduce_kernel<<<ew_grid(n), 256, 0, s>>>(a.ws, p->dwp, n, w.splits);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_wgrad_fold_complex(const float* dwp, int ntaps, int cin, int cout, int transposed, float* dw_r, float* dw_i, void* stream) {
  DCS_REQUIRE(dwp && dw_r && dw_i && ntaps > 0 && cin > 0 && cout > 0, "dcs_wgrad_fold_complex: bad arguments");
  wgrad_fold_complex_kernel<<<ew_grid((int64_t)ntaps * cin * cout), 256, 0, (cudaStream_t)stream>>>(dwp, ntaps, cin, cout, transposed, dw_r, dw_i);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_transpose(const float* src, float* dst, int rows, int cols, int src_pitch, void* stream) {
  DCS_REQUIRE(src && dst && rows > 0 && cols > 0 && src_pitch >= cols, "dcs_transpose: bad arguments");
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, rows, cols, src_pitch);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_sgemm(const float* A, int lda, const float* B, int ldb_n, int ldb_k, const float* bias, float* C, int ldc, int M, int N, int K,
                         int accumulate, void* stream) {
  DCS_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && lda >= K && ldc >= N, "dcs_sgemm: bad arguments");
  sgemm_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb_n, ldb_k, bias, C, ldc, M, N, K, accumulate);
  DCS_LAUNCHED();
  return 0;
}

static void colsum_chunking(int64_t rows, int* n_chunks, int64_t* rpc) {
  int64_t ctas = std::max<int64_t>(1, std::min<int64_t>((rows + 63) / 64, 2 * (int64_t)num_sms()));
  *rpc = (rows + ctas - 1) / ctas;
  *n_chunks = (int)((rows + *rpc - 1) / *rpc);
}
extern "C" int64_t dcs_colsum_workspace_bytes(int64_t rows, int cols) {
  if (rows <= 0 || cols <= 0) return -1;
  int nc; int64_t rpc;
  colsum_chunking(rows, &nc, &rpc);
  return (int64_t)nc * cols * sizeof(double);
}
extern "C" int dcs_colsum(const float* x, int64_t rows, int cols, int pitch, int mode, float* out0, float* out1, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  DCS_REQUIRE(x && out0 && workspace && rows > 0 && cols > 0 && pitch >= cols && (mode == 0 || (mode == 1 && out1 && cols % 2 == 0)),
              "dcs_colsum: bad arguments");
  DCS_REQUIRE(workspace_bytes >= dcs_colsum_workspace_bytes(rows, cols), "dcs_colsum: workspace too small");
  int nc; int64_t rpc;
  colsum_chunking(rows, &nc, &rpc);
  cudaStream_t s = (cudaStream_t)stream;
  colsum_kernel<<<nc, 256, 0, s>>>(x, rows, cols, pitch, rpc, reinterpret_cast<double*>(workspace));
  DCS_LAUNCHED();
  colsum_finalize_kernel<<<(cols + 7) / 8, 256, 0, s>>>(reinterpret_cast<const double*>(workspace), nc, cols, mode, out0, out1);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_dilate(const float* dy, float* out, int batch, int out_h, int out_w, int in_h, int in_w, int channels, int stride_h,
                          int stride_w, void* stream) {
  DCS_REQUIRE(dy && out && batch > 0 && out_h > 0 && out_w > 0 && in_h > 0 && in_w > 0 && channels > 0 && stride_h > 0 && stride_w > 0 &&
              (out_h - 1) * stride_h < in_h && (out_w - 1) * stride_w < in_w, "dcs_dilate: bad arguments");
  const int64_t n = (int64_t)batch * in_h * in_w * channels;
  dilate_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const float2*)dy, (float2*)out, batch, out_h, out_w, in_h, in_w, channels, stride_h, stride_w);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_cconv_dgrad_cin1(const float* dy, const float* w_r, const float* w_i, float* dx, int batch, int in_h, int in_w, int out_h, int out_w,
                                    int cout, int kh, int kw, int stride_h, int stride_w, void* stream) {
  DCS_REQUIRE(dy && w_r && w_i && dx && batch > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && cout > 0 && kh > 0 && kw > 0 && (kh & 1) && (kw & 1) &&
              stride_h > 0 && stride_w > 0 && kh * kw * cout <= 4096, "dcs_cconv_dgrad_cin1: bad arguments");
  const int64_t n = (int64_t)batch * in_h * in_w;
  cconv_dgrad_cin1_kernel<<<ew_grid(n), 256, (size_t)kh * kw * cout * sizeof(float2), (cudaStream_t)stream>>>(
      (const float2*)dy, w_r, w_i, (float2*)dx, batch, in_h, in_w, out_h, out_w, cout, kh, kw, stride_h, stride_w);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_upcat_fwd(const void* d, const void* skip, int in_dtype, void* z, int out_dtype, int batch, int h, int w, int c0, int c1, int up_h, int up_w,
                             void* stream) {
  DCS_REQUIRE(d && z && (skip || c1 == 0) && batch > 0 && h > 0 && w > 0 && c0 > 0 && c1 >= 0 && up_h >= 1 && up_w >= 1 && is_dtype(out_dtype) && is_dtype(in_dtype),
              "dcs_upcat_fwd: bad arguments");
  const int64_t n = (int64_t)batch * h * up_h * w * up_w * (c0 + c1);
  cudaStream_t s = (cudaStream_t)stream;
  if ((c0 & 1) == 0 && (c1 & 1) == 0 && (((uintptr_t)d | (uintptr_t)skip) & 15) == 0) {
    const int64_t n2 = n / 2;
    if (out_dtype == DCS_F32) upcat_fwd2_kernel<float><<<ew_grid(n2), 256, 0, s>>>(d, skip, in_dtype, (float*)z, batch, h, w, c0 / 2, c1 / 2, up_h, up_w);
    else if (out_dtype == DCS_F16) upcat_fwd2_kernel<__half><<<ew_grid(n2), 256, 0, s>>>(d, skip, in_dtype, (__half*)z, batch, h, w, c0 / 2, c1 / 2, up_h, up_w);
    else upcat_fwd2_kernel<__nv_bfloat16><<<ew_grid(n2), 256, 0, s>>>(d, skip, in_dtype, (__nv_bfloat16*)z, batch, h, w, c0 / 2, c1 / 2, up_h, up_w);
    DCS_LAUNCHED();
    return 0;
  }
  if (out_dtype == DCS_F32) upcat_fwd_kernel<float><<<ew_grid(n), 256, 0, s>>>(d, skip, in_dtype, (float*)z, batch, h, w, c0, c1, up_h, up_w);
  else if (out_dtype == DCS_F16) upcat_fwd_kernel<__half><<<ew_grid(n), 256, 0, s>>>(d, skip, in_dtype, (__half*)z, batch, h, w, c0, c1, up_h, up_w);
  else upcat_fwd_kernel<__nv_bfloat16><<<ew_grid(n), 256, 0, s>>>(d, skip, in_dtype, (__nv_bfloat16*)z, batch, h, w, c0, c1, up_h, up_w);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int64_t dcs_dec6_bwd_workspace_bytes(void) { return (int64_t)2 * num_sms() * (2 * 9 * kD6C + 2) * (int64_t)sizeof(double); }
extern "C" int dcs_dec6_bwd(const void* d, const void* skip, int in_dtype, const float* dpre, const float* w_r, const float* w_i, int batch, int h, int w, int c0, int c1,
                            float* g_d, float* g_skip, float* dw_r, float* dw_i, float* db_r, float* db_i, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  DCS_REQUIRE(d && skip && dpre && w_r && w_i && g_d && g_skip && dw_r && dw_i && db_r && db_i && workspace, "dcs_dec6_bwd: null pointer");
  DCS_REQUIRE(batch > 0 && h > 0 && w > 0 && c0 > 0 && c1 > 0 && c0 + c1 <= kD6C && is_dtype(in_dtype), "dcs_dec6_bwd: c0 + c1 must be <= 16");
  DCS_REQUIRE(workspace_bytes >= dcs_dec6_bwd_workspace_bytes(), "dcs_dec6_bwd: workspace too small");
  const int n_tiles = batch * ((h + kD6TH - 1) / kD6TH) * ((w + kD6TW - 1) / kD6TW);
  const int ctas = std::min(n_tiles, 2 * num_sms());
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = sizeof(float2) * (9 * kD6TH * kD6TW + kD6TH * kD6TW * (kD6C + 1));
  DCS_CUDA(cudaFuncSetAttribute(dec6_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dec6_bwd_kernel<<<ctas, 256, smem, s>>>(d, skip, in_dtype, (const float2*)dpre, w_r, w_i, batch, h, w, c0, c1, (float2*)g_d,
                                       (float2*)g_skip, reinterpret_cast<double*>(workspace));
  DCS_LAUNCHED();
  dec6_bwd_finalize_kernel<<<1, 160, 0, s>>>(reinterpret_cast<const double*>(workspace), ctas, c0 + c1, dw_r, dw_i, db_r, db_i);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_act_bwd(const void* y, int y_dtype, const float* g0, const float* g1, const float* chan_const, float* dz, int batch, int64_t hw,
                           int channels, int act, void* stream) {
  DCS_REQUIRE(g0 && dz && (y || act == DCS_ACT_NONE) && batch > 0 && hw > 0 && channels > 0 && act >= DCS_ACT_NONE && act <= DCS_ACT_LRELU &&
              is_dtype(y_dtype), "dcs_act_bwd: bad arguments");
  const int64_t n = (int64_t)batch * hw * channels;
  act_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(y, y_dtype, (const float2*)g0, (const float2*)g1, (const float2*)chan_const,
                                                               (float2*)dz, hw * channels, channels, n, act);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_dropout(const void* x, void* y, int64_t n, int dtype, float p, uint64_t seed, uint64_t offset, void* stream) {
  DCS_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f && is_dtype(dtype), "dcs_dropout: bad arguments");
  const float sc = 1.f / (1.f - p);
  if (dtype == DCS_F32) dropout_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, n, p, sc, seed, offset);
  else if (dtype == DCS_F16) dropout_h16_kernel<__half><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, n, p, sc, seed, offset);
  else dropout_h16_kernel<__nv_bfloat16><<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, p, sc, seed, offset);
  DCS_LAUNCHED();
  return 0;
}

static int att_G(int C) { return C >= 32 ? 32 : C; }
static int att_chunks(int hw, int G) { return (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)hw + (256 / G) * 8 - 1) / ((256 / G) * 8), 64)); }

extern "C" int64_t dcs_attention_bwd_workspace_bytes(int batch, int h, int w, int channels, int reduced) {
  if (batch <= 0 || h <= 0 || w <= 0 || channels <= 0 || channels > 256 || (channels & (channels - 1)) || reduced <= 0 || reduced > 16) return -1;
  const int G = att_G(channels), chunks = att_chunks(h * w, G);
  const int64_t tiles = (int64_t)batch * ((h + kW7TH - 1) / kW7TH) * ((w + kW7TW - 1) / kW7TW);
  return (int64_t)batch * chunks * channels * 2 * sizeof(double) + tiles * 196 * sizeof(double) + (int64_t)(batch + 1) * 4 * reduced * channels * sizeof(float) +
         (int64_t)batch * h * w * 4 * sizeof(float) + 16;
}

extern "C" int dcs_attention_bwd(const dcs_attention_bwd_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->dy && p->gate_c && p->stats && p->gate_s && p->w7 && p->sums && p->w1_r && p->w1_i && p->w2_r && p->w2_i && p->dspre &&
              p->dx && p->chan_const && p->dw1_r && p->dw1_i && p->dw2_r && p->dw2_i && p->dw7_r && p->dw7_i && p->workspace,
              "dcs_attention_bwd: null pointer");
  const int64_t need = dcs_attention_bwd_workspace_bytes(p->batch, p->h, p->w, p->channels, p->reduced);
  DCS_REQUIRE(need > 0, "dcs_attention_bwd: channels must be a power of two <= 256, reduced <= 16");
  DCS_REQUIRE(is_dtype(p->x_dtype), "dcs_attention_bwd: bad x_dtype");
  DCS_REQUIRE(p->workspace_bytes >= need, "dcs_attention_bwd: workspace too small");
  const int C = p->channels, R = p->reduced, hw = p->h * p->w, G = att_G(C), chunks = att_chunks(hw, G);
  const int tiles_y = (p->h + kW7TH - 1) / kW7TH, tiles_x = (p->w + kW7TW - 1) / kW7TW;
  const int n_tiles = p->batch * tiles_y * tiles_x;
  double* da_partial = reinterpret_cast<double*>(p->workspace);
  double* w7_partial = da_partial + (int64_t)p->batch * chunks * C * 2;
  float* wpart = reinterpret_cast<float*>(w7_partial + (int64_t)n_tiles * 196);
  cudaStream_t s = (cudaStream_t)stream;
  const void* x = p->x;
  const int xdt = p->x_dtype;
  const float2 *dy = (const float2*)p->dy, *gc = (const float2*)p->gate_c, *gsp = (const float2*)p->gate_s;
  // (the thread-per-pixel kernels use 16-byte accesses: fall back to the lane-per-channel kernels for unaligned tensors)
  const bool small = (C == 8 || C == 16) && !getenv("DCS_ATT_BWD_NO_SMALL") && ((((uintptr_t)p->x | (uintptr_t)p->dy | (uintptr_t)p->dx) & 15) == 0);
  (void)x;
  if (small && C == 8) att_bwd_ds_small_kernel<8><<<dim3(chunks, p->batch), 256, 0, s>>>(x, xdt, (const float4*)p->dy, gc, gsp, (float2*)p->dspre, hw);
  else if (small) att_bwd_ds_small_kernel<16><<<dim3(chunks, p->batch), 256, 0, s>>>(x, xdt, (const float4*)p->dy, gc, gsp, (float2*)p->dspre, hw);
  else att_bwd_ds_kernel<<<dim3(chunks, p->batch), 256, 0, s>>>(x, xdt, dy, gc, gsp, (float2*)p->dspre, hw, C, G);
  DCS_LAUNCHED();
  att_bwd_w7_kernel<<<n_tiles, 256, 0, s>>>((const float4*)p->stats, (const float2*)p->dspre, p->h, p->w, tiles_x, tiles_y, w7_partial);
  DCS_LAUNCHED();
  att_bwd_w7_finalize_kernel<<<196, 256, 0, s>>>(w7_partial, n_tiles, p->dw7_r, p->dw7_i);
  DCS_LAUNCHED();
  float4* dstats = reinterpret_cast<float4*>(((uintptr_t)(wpart + (int64_t)(p->batch + 1) * 4 * R * C) + 15) & ~(uintptr_t)15);
  att_bwd_dstats_kernel<<<dim3((p->w + kDsTW - 1) / kDsTW, (p->h + kDsTH - 1) / kDsTH, p->batch), 256, 0, s>>>((const float2*)p->dspre, p->w7, dstats, p->h, p->w);
  DCS_LAUNCHED();
  if (small && C == 8) att_bwd_du_small_kernel<8><<<dim3(chunks, p->batch), 256, 0, s>>>(x, xdt, (const float4*)p->dy, gc, gsp, dstats, (float4*)p->dx, da_partial, hw);
  else if (small) att_bwd_du_small_kernel<16><<<dim3(chunks, p->batch), 256, 0, s>>>(x, xdt, (const float4*)p->dy, gc, gsp, dstats, (float4*)p->dx, da_partial, hw);
  else att_bwd_du_kernel<<<dim3(chunks, p->batch), 256, 0, s>>>(x, xdt, dy, gc, gsp, dstats, (float2*)p->dx, da_partial, p->h, p->w, C, G);
  DCS_LAUNCHED();
  att_bwd_gate_kernel<<<p->batch, 256, 0, s>>>(da_partial, chunks, gc, (const long long*)p->sums, 1.f / (float)hw, C, R, p->w1_r, p->w1_i, p->w2_r,
                                               p->w2_i, (float2*)p->chan_const, wpart);
  DCS_LAUNCHED();
  const int rc = R * C;
  float* red = wpart + (int64_t)p->batch * 4 * rc;
  batch_reduce_kernel<<<(4 * rc + 255) / 256, 256, 0, s>>>(wpart, red, p->batch, 4 * rc);
  DCS_LAUNCHED();
  DCS_CUDA(cudaMemcpyAsync(p->dw1_r, red, rc * sizeof(float), cudaMemcpyDeviceToDevice, s));
  DCS_CUDA(cudaMemcpyAsync(p->dw1_i, red + rc, rc * sizeof(float), cudaMemcpyDeviceToDevice, s));
  DCS_CUDA(cudaMemcpyAsync(p->dw2_r, red + 2 * rc, rc * sizeof(float), cudaMemcpyDeviceToDevice, s));
  DCS_CUDA(cudaMemcpyAsync(p->dw2_i, red + 3 * rc, rc * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int dcs_lstm_train_fwd(const float* pre, const float* w_hh, int n_seq, int n_groups, int steps, int hidden, float* h, float* gates,
                                  float* cells, void* stream) {
  DCS_REQUIRE(pre && w_hh && h && gates && cells && n_seq > 0 && n_groups > 0 && n_seq % n_groups == 0 && steps > 0, "dcs_lstm_train_fwd: bad arguments");
  DCS_REQUIRE(hidden == 64, "dcs_lstm_train_fwd: only hidden = 64 is built");
  lstm_train_fwd_kernel<64><<<dim3(n_seq, 2), 256, 0, (cudaStream_t)stream>>>(pre, w_hh, n_seq, n_groups, steps, h, gates, cells);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_lstm_train_bwd(const float* w_hh, const float* gates, const float* cells, const float* dh, int n_seq, int n_groups, int steps,
                                  int hidden, float* dpre, void* stream) {
  DCS_REQUIRE(w_hh && gates && cells && dh && dpre && n_seq > 0 && n_groups > 0 && n_seq % n_groups == 0 && steps > 0, "dcs_lstm_train_bwd: bad arguments");
  DCS_REQUIRE(hidden == 64, "dcs_lstm_train_bwd: only hidden = 64 is built");
  lstm_train_bwd_kernel<64><<<dim3(n_seq, 2), 256, 0, (cudaStream_t)stream>>>(w_hh, gates, cells, dh, n_seq, n_groups, steps, dpre);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_cplx_split(const void* x, int in_dtype, float* planes, int64_t n, void* stream) {
  DCS_REQUIRE(x && planes && n > 0 && is_dtype(in_dtype), "dcs_cplx_split: bad arguments");
  cplx_split_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, in_dtype, planes, n);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_cplx_merge(const float* planes, float* x, int64_t n, void* stream) {
  DCS_REQUIRE(x && planes && n > 0, "dcs_cplx_merge: bad arguments");
  cplx_merge_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(planes, (float2*)x, n);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_clstm_combine(const float* h, float* out, int64_t n, void* stream) {
  DCS_REQUIRE(h && out && n > 0, "dcs_clstm_combine: bad arguments");
  clstm_combine_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(h, (float2*)out, n);
  DCS_LAUNCHED();
  return 0;
}
extern "C" int dcs_clstm_combine_bwd(const float* dout, float* dh, int64_t n, void* stream) {
  DCS_REQUIRE(dh && dout && n > 0, "dcs_clstm_combine_bwd: bad arguments");
  clstm_combine_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const float2*)dout, dh, n);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_sumsq(const float* x, int64_t n, double* out, int accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
  const int grid = 2 * num_sms();
  DCS_REQUIRE(x && out && workspace && n > 0 && workspace_bytes >= (int64_t)grid * (int64_t)sizeof(double), "dcs_sumsq: bad arguments (workspace >= 8 * 2 * SMs bytes)");
  sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, reinterpret_cast<double*>(workspace));
  DCS_LAUNCHED();
  sumsq_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(workspace), grid, out, accumulate);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_adam_amsgrad(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int step, const double* grad_sumsq, float max_norm,
                                float grad_scale, void* stream) {
  DCS_REQUIRE(param && grad && exp_avg && exp_avg_sq && max_exp_avg_sq && n > 0 && step >= 1, "dcs_adam_amsgrad: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_amsgrad_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                    weight_decay, bc1, bc2s, grad_sumsq, max_norm, grad_scale);
  DCS_LAUNCHED();
  return 0;
}

extern "C" int dcs_gather_pack(const float* src, const int32_t* idx4, const int8_t* sign4, void* dst, int64_t n, int out_dtype, void* stream) {
  DCS_REQUIRE(src && idx4 && sign4 && dst && n > 0 && (is_dtype(out_dtype) || out_dtype == 3), "dcs_gather_pack: bad arguments");
  gather_pack_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(src, (const int4*)idx4, (const char4*)sign4, dst, n, out_dtype);
  DCS_LAUNCHED();
  return 0;
}
