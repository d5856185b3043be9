// lstm.cu — ComplexLSTM latent (/root/reference/c_network.py:12-51, built at 117-123; no cuDNN).
//
// real_lstm / imag_lstm = nn.LSTM(128 -> 64, num_layers=2, bidirectional, batch_first); the reference runs four
// passes R(re), I(re), R(im), I(im) and combines out = (R(re) - I(im)) + j (R(im) + I(re))  (c_network.py:38-43).
//
// Here the four passes are batched as 4*B independent sequences q = (lstm, part, b).  Per layer:
//   (1) input projection for all time steps as ONE real GEMM on the implicit-GEMM convolution kernel (1x1 tap),
//   (2) a persistent recurrent kernel: a CTA owns NSEQ sequences of one (lstm, direction); each of its 256 threads
//       keeps one row of W_hh (64 floats) in registers, h lives in shared memory, the pre-activations of step t+1
//       are prefetched while step t is computed.  Gate order i, f, g, o (torch.nn.LSTM).
// The recurrence is latency-bound (S = 2*T/8 sequential steps); it is reported separately from both rooflines.
#include <string.h>
#include <algorithm>
#include "common.cuh"

namespace dcs {

constexpr int kH = 64;        // hidden size (channels[4] // 2)
constexpr int kG = 4 * kH;    // gate rows

template <typename T>
__global__ void lstm_deinterleave_kernel(const T* __restrict__ x, float* __restrict__ xp, int64_t rows, int D) {
  // x: (rows, D) complex  ->  xp[part][rows][D]
  const int64_t n = rows * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = Elem<T>::ldc(x, i);
    xp[i] = v.x;
    xp[n + i] = v.y;
  }
}

// pre: gate pre-activations (bias included).  Row (part*B + b)*S + t with row stride `ld`; the (lstm, dir) block
// starts at pre0 + lstm*stride_lstm + dir*stride_dir.  hout: [q][S][2*kH], this direction writes columns dir*kH...
template <int NSEQ>
__global__ void __launch_bounds__(256, 2) lstm_recurrent_kernel(const float* __restrict__ pre0, int64_t stride_lstm,
                                                                int64_t stride_dir, int ld, const float* __restrict__ whh,
                                                                float* __restrict__ hout, int B, int S) {
  // grid.x = (4B / NSEQ), grid.y = dir
  __shared__ __align__(16) float hs[NSEQ][kH];
  __shared__ float gs[NSEQ][kG];
  const int tid = threadIdx.x;
  const int dir = blockIdx.y;
  const int q0 = blockIdx.x * NSEQ;
  const int lstm = q0 / (2 * B);
  const float* wrow = whh + ((int64_t)(lstm * 2 + dir) * kG + tid) * kH;
  float w[kH];
#pragma unroll
  for (int k = 0; k < kH; k += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(wrow + k));
    w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
  }
  const float* pre = pre0 + lstm * stride_lstm + dir * stride_dir + tid;
  int64_t prow[NSEQ];
#pragma unroll
  for (int s = 0; s < NSEQ; ++s) {
    const int q = q0 + s;
    const int pb = q % (2 * B);  // part*B + b
    prow[s] = (int64_t)pb * S;
  }
  // cell role: thread -> (sequence cs, unit cu)
  const int cs = tid / kH, cu = tid % kH;
  float c_state = 0.f;
  for (int i = tid; i < NSEQ * kH; i += 256) (&hs[0][0])[i] = 0.f;
  // register prefetch queue: the pre-activations of steps t+1 .. t+kPF are in flight while step t is computed
  // (an HBM round trip is ~2-3 recurrence steps long)
  constexpr int kPF = 4;
  float pq[kPF][NSEQ];
#pragma unroll
  for (int d = 0; d < kPF; ++d) {
    const int st = min(d, S - 1);
    const int t = dir ? S - 1 - st : st;
#pragma unroll
    for (int s = 0; s < NSEQ; ++s) pq[d][s] = __ldg(pre + (prow[s] + t) * ld);
  }
  __syncthreads();
  for (int step = 0; step < S; ++step) {
    const int t = dir ? S - 1 - step : step;
    float acc[NSEQ], acc2[NSEQ];  // two partial sums per sequence: halves the dependent-FMA chain length
#pragma unroll
    for (int s = 0; s < NSEQ; ++s) { acc[s] = pq[0][s]; acc2[s] = 0.f; }
#pragma unroll
    for (int d = 0; d + 1 < kPF; ++d)
#pragma unroll
      for (int s = 0; s < NSEQ; ++s) pq[d][s] = pq[d + 1][s];
    {
      const int st = min(step + kPF, S - 1);
      const int tn = dir ? S - 1 - st : st;
#pragma unroll
      for (int s = 0; s < NSEQ; ++s) pq[kPF - 1][s] = __ldg(pre + (prow[s] + tn) * ld);
    }
#pragma unroll
    for (int k = 0; k < kH; k += 4) {
#pragma unroll
      for (int s = 0; s < NSEQ; ++s) {
        const float4 h4 = *reinterpret_cast<const float4*>(&hs[s][k]);
        acc[s] = fmaf(w[k], h4.x, acc[s]);
        acc2[s] = fmaf(w[k + 1], h4.y, acc2[s]);
        acc[s] = fmaf(w[k + 2], h4.z, acc[s]);
        acc2[s] = fmaf(w[k + 3], h4.w, acc2[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < NSEQ; ++s) gs[s][tid] = acc[s] + acc2[s];
    __syncthreads();
    if (cs < NSEQ) {
      const float ig = fast_sigmoid(gs[cs][cu]);
      const float fg = fast_sigmoid(gs[cs][kH + cu]);
      const float gg = fast_tanh(gs[cs][2 * kH + cu]);
      const float og = fast_sigmoid(gs[cs][3 * kH + cu]);
      c_state = fg * c_state + ig * gg;
      const float h = og * fast_tanh(c_state);
      hs[cs][cu] = h;
      hout[((int64_t)(q0 + cs) * S + t) * (2 * kH) + dir * kH + cu] = h;
    }
    __syncthreads();
  }
}

__global__ void lstm_combine_kernel(const float* __restrict__ h1, float2* __restrict__ y, int64_t n_per_q4) {
  // h1: [lstm][part][B*S*2H];  out.re = R(re) - I(im), out.im = R(im) + I(re)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_per_q4; i += (int64_t)gridDim.x * blockDim.x) {
    const float r_re = h1[i], r_im = h1[n_per_q4 + i], i_re = h1[2 * n_per_q4 + i], i_im = h1[3 * n_per_q4 + i];
    y[i] = make_float2(r_re - i_im, r_im + i_re);
  }
}

struct LstmWs {
  float *xp, *pre, *h0, *h1;
  int64_t total;
};
static LstmWs carve(void* ws, int B, int S, int D) {
  LstmWs w;
  const int64_t rows = (int64_t)B * S;
  float* p = reinterpret_cast<float*>(ws);
  w.xp = p; p += 2 * rows * D;
  w.pre = p; p += 2 * rows * 4 * kG;         // layer 0: [2BS][1024]; layer 1: [2 lstm][2BS][512] (same size)
  w.h0 = p; p += 4 * rows * 2 * kH;
  w.h1 = p; p += 4 * rows * 2 * kH;
  w.total = (p - reinterpret_cast<float*>(ws)) * (int64_t)sizeof(float);
  return w;
}

static int gemm_rows(const float* a, int64_t rows, int K, const float* w, const float* bias, int N, float* out, void* stream) {
  dcs_cconv_params c;
  memset(&c, 0, sizeof(c));
  c.src0 = a; c.c0 = K / 2; c.c1 = 0;
  c.batch = 1; c.in_h = 1; c.in_w = (int)rows; c.out_h = 1; c.out_w = (int)rows; c.cout = N / 2;
  c.up_h = c.up_w = 1; c.stride_h = c.stride_w = 1; c.ntaps = 1;
  c.weight = w; c.bias = bias; c.act = DCS_ACT_NONE; c.dst = out; c.in_dtype = DCS_F32; c.out_dtype = DCS_F32;
  return dcs_cconv2d_fwd(&c, stream);
}

static void launch_rec(int nseq, const float* pre, int64_t stride_lstm, int64_t stride_dir, int ld, const float* whh,
                       float* hout, int B, int S, cudaStream_t s) {
  dim3 grid(4 * B / nseq, 2);
  if (nseq == 4) lstm_recurrent_kernel<4><<<grid, 256, 0, s>>>(pre, stride_lstm, stride_dir, ld, whh, hout, B, S);
  else lstm_recurrent_kernel<2><<<grid, 256, 0, s>>>(pre, stride_lstm, stride_dir, ld, whh, hout, B, S);
}

// tensor-core input projection (tf32 operands read from fp32 memory): out[rows][256] = a[rows][K] * w[256][K]^T + bias
static int gemm_rows_tc(const float* a, int64_t rows, int K, const float* w_t, const float* bias, float* out, void* stream) {
  dcs_cconv_params c;
  memset(&c, 0, sizeof(c));
  c.src0 = a; c.c0 = K / 2; c.c1 = 0;
  c.batch = 1; c.in_h = 1; c.in_w = (int)rows; c.out_h = 1; c.out_w = (int)rows; c.cout = kG / 2;
  c.up_h = c.up_w = 1; c.stride_h = c.stride_w = 1; c.ntaps = 1;
  c.weight = w_t; c.bias = bias; c.act = DCS_ACT_NONE; c.dst = out; c.in_dtype = DCS_F32; c.out_dtype = DCS_F32;
  return dcs_cconv2d_tc_fwd(&c, stream);
}

}  // namespace dcs

using namespace dcs;

extern "C" int64_t dcs_clstm_workspace_bytes(int batch, int seq, int hidden) {
  if (batch <= 0 || seq <= 0 || hidden != kH) return -1;
  return carve(nullptr, batch, seq, 2 * kH).total;
}

extern "C" int dcs_clstm_fwd(const dcs_clstm_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->w_ih0 && p->w_ih1 && p->w_hh && p->bias && p->workspace, "dcs_clstm_fwd: null pointer");
  DCS_REQUIRE(p->hidden == kH && p->in_dim == 2 * kH, "dcs_clstm_fwd: only ComplexLSTM(128 -> 64) is built (got %d -> %d)", p->in_dim, p->hidden);
  DCS_REQUIRE(p->batch > 0 && p->seq > 0, "dcs_clstm_fwd: bad shape");
  const int B = p->batch, S = p->seq, D = p->in_dim;
  LstmWs w = carve(p->workspace, B, S, D);
  DCS_REQUIRE(p->workspace_bytes >= w.total, "dcs_clstm_fwd: workspace too small (%lld < %lld)", (long long)p->workspace_bytes, (long long)w.total);
  DCS_REQUIRE(2ll * B * S < (1ll << 31), "dcs_clstm_fwd: batch*seq too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * S;
  {
    const int64_t n = rows * D;
    const int g = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
    if (p->in_dtype == DCS_BF16) lstm_deinterleave_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)p->x, w.xp, rows, D);
    else lstm_deinterleave_kernel<float><<<g, 256, 0, s>>>((const float*)p->x, w.xp, rows, D);
    DCS_LAUNCHED();
  }
  // NSEQ must divide 2B; 2 sequences per CTA (two co-resident CTAs per SM) unless that overflows the machine
  int nseq = 2;
  if ((2 * B) % 4 == 0 && (4 * B / 2) * 2 > 2 * num_sms()) nseq = 4;
  if (p->seqs_per_cta == 4 && (2 * B) % 4 == 0) nseq = 4;  // one CTA per SM: leaves room for kernels on other streams
  if (p->seqs_per_cta == 2) nseq = 2;
  const float* whh1 = p->w_hh + (int64_t)4 * kG * kH;
  const int64_t rows2 = 2 * rows;
  if (p->w_ih0_t && p->w_ih1_t) {
    // ---- tensor-core (tf32) input projections: pre[(lstm,dir)][(part,b,s)][4H], eight N=256 GEMMs
    for (int sl = 0; sl < 4; ++sl)
      if (int e = gemm_rows_tc(w.xp, rows2, D, p->w_ih0_t + (int64_t)sl * kG * D, p->bias + sl * kG, w.pre + sl * rows2 * kG, stream)) return e;
    launch_rec(nseq, w.pre, 2 * rows2 * kG, rows2 * kG, kG, p->w_hh, w.h0, B, S, s);
    DCS_LAUNCHED();
    for (int sl = 0; sl < 4; ++sl)
      if (int e = gemm_rows_tc(w.h0 + (int64_t)(sl / 2) * rows2 * 2 * kH, rows2, 2 * kH, p->w_ih1_t + (int64_t)sl * kG * 2 * kH,
                               p->bias + 4 * kG + sl * kG, w.pre + sl * rows2 * kG, stream)) return e;
    launch_rec(nseq, w.pre, 2 * rows2 * kG, rows2 * kG, kG, whh1, w.h1, B, S, s);
    DCS_LAUNCHED();
  } else {
    // ---- fp32 CUDA-core projections.  layer 0: pre[(part,b,s)][lstm][dir][4H]
    if (int e = gemm_rows(w.xp, rows2, D, p->w_ih0, p->bias, 4 * kG, w.pre, stream)) return e;
    launch_rec(nseq, w.pre, 2 * kG, kG, 4 * kG, p->w_hh, w.h0, B, S, s);
    DCS_LAUNCHED();
    // layer 1: per lstm, rows (part,b,s) of h0[lstm] -> pre1[lstm][(part,b,s)][dir][4H]
    for (int l = 0; l < 2; ++l) {
      if (int e = gemm_rows(w.h0 + (int64_t)l * rows2 * 2 * kH, rows2, 2 * kH, p->w_ih1 + (int64_t)l * 2 * kH * 2 * kG,
                            p->bias + 4 * kG + l * 2 * kG, 2 * kG, w.pre + (int64_t)l * rows2 * 2 * kG, stream))
        return e;
    }
    launch_rec(nseq, w.pre, rows2 * 2 * kG, kG, 2 * kG, whh1, w.h1, B, S, s);
    DCS_LAUNCHED();
  }
  {
    const int64_t n = rows * 2 * kH;
    const int g = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
    lstm_combine_kernel<<<g, 256, 0, s>>>(w.h1, (float2*)p->y, n);
    DCS_LAUNCHED();
  }
  return 0;
}
