// rlstm.cu — the real network's latent LSTM (/root/reference/r_network.py:70-74, 137-139):
// nn.LSTM(256 -> 128, num_layers=2, bidirectional, batch_first) over the flattened latent (sequence index = h * W + w,
// features = channels: exactly the channels-last tensor (B, H, W, C) viewed as (B, S, C)).
// First, straightforward fp32 version: per layer ONE fp32 GEMM for the input projections of both directions
// (dcs_cconv2d_fwd as a 1x1 conv, bias = b_ih + b_hh) and a recurrent kernel with one CTA per (sequence, direction),
// one thread per gate row, W_hh^T streamed from L2 every step (coalesced across the gate rows).  Exact expf / tanhf.
// The tensor-core recurrence of lstm.cu (register-resident W_hh fragments) is the follow-up for hidden = 128.
#include <string.h>
#include "common.cuh"

namespace dcs {

// pre: [b * S + t][ld] gate pre-activations (input projection + both biases), this direction's 4H columns at dir * 4H,
// gate order i, f, g, o.  whh_t: [dir][H][4H].  hout: [b * S + t][2H], this direction at dir * H.
__global__ void __launch_bounds__(512) rlstm_recurrent_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ whh_t,
                                                              float* __restrict__ hout, int S, int H) {
  extern __shared__ float rl_smem[];
  float* h = rl_smem;            // [H]
  float* gates = rl_smem + H;    // [4H]
  const int b = blockIdx.x, dir = blockIdx.y, r = threadIdx.x, G4 = 4 * H;
  const float* wt = whh_t + (int64_t)dir * H * G4 + r;
  const float* prow = pre + (int64_t)b * S * ld + dir * G4 + r;
  float* hrow = hout + (int64_t)b * S * 2 * H + dir * H + r;
  if (r < H) h[r] = 0.f;
  float c = 0.f;
  __syncthreads();
  for (int step = 0; step < S; ++step) {
    const int t = dir ? S - 1 - step : step;
    float g = prow[(int64_t)t * ld];
#pragma unroll 8
    for (int k = 0; k < H; ++k) g = fmaf(__ldg(wt + (int64_t)k * G4), h[k], g);
    gates[r] = g;
    __syncthreads();
    float hn = 0.f;
    if (r < H) {
      const float ig = sigmoidf_(gates[r]), fg = sigmoidf_(gates[H + r]), gg = tanhf(gates[2 * H + r]), og = sigmoidf_(gates[3 * H + r]);
      c = fg * c + ig * gg;
      hn = og * tanhf(c);
      hrow[(int64_t)t * 2 * H] = hn;
    }
    __syncthreads();             // every thread has read the old h
    if (r < H) h[r] = hn;
    __syncthreads();
  }
}

}  // namespace dcs

using namespace dcs;

static int rl_gemm_rows(const float* a, int64_t rows, int K, const float* w, const float* bias, int N, float* out, void* stream) {
  dcs_cconv_params c;
  memset(&c, 0, sizeof(c));
  c.src0 = a; c.c0 = K / 2; c.c1 = 0;
  c.batch = 1; c.in_h = 1; c.in_w = (int)rows; c.out_h = 1; c.out_w = (int)rows; c.cout = N / 2;
  c.up_h = c.up_w = 1; c.stride_h = c.stride_w = 1; c.ntaps = 1;
  c.weight = w; c.bias = bias; c.act = DCS_ACT_NONE; c.dst = out; c.in_dtype = DCS_F32; c.out_dtype = DCS_F32;
  return dcs_cconv2d_fwd(&c, stream);
}

extern "C" int64_t dcs_rlstm_workspace_bytes(int batch, int seq, int hidden) {
  if (batch <= 0 || seq <= 0 || hidden <= 0) return -1;
  return (int64_t)batch * seq * (8 * hidden + 2 * hidden) * 4;     // pre [BS][8H] + layer-0 output [BS][2H]
}

extern "C" int dcs_rlstm_fwd(const dcs_rlstm_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->w_ih0_t && p->w_ih1_t && p->w_hh_t && p->bias && p->workspace, "dcs_rlstm_fwd: null pointer");
  DCS_REQUIRE(p->in_dtype == DCS_F32, "dcs_rlstm_fwd: fp32 input only in this version");
  const int B = p->batch, S = p->seq, D = p->in_dim, H = p->hidden;
  DCS_REQUIRE(B > 0 && B <= 65535 && S > 0 && H == 128 && D % 2 == 0 && D > 0, "dcs_rlstm_fwd: only nn.LSTM(D -> 128) is built (got %d -> %d)", D, H);
  DCS_REQUIRE(p->workspace_bytes >= dcs_rlstm_workspace_bytes(B, S, H), "dcs_rlstm_fwd: workspace too small");
  DCS_REQUIRE((int64_t)B * S < (1ll << 31), "dcs_rlstm_fwd: batch*seq too large");
  float* pre = reinterpret_cast<float*>(p->workspace);
  float* h0 = pre + (int64_t)B * S * 8 * H;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * S;
  const size_t smem = (size_t)5 * H * sizeof(float);
  // layer 0
  if (int e = rl_gemm_rows((const float*)p->x, rows, D, p->w_ih0_t, p->bias, 8 * H, pre, stream)) return e;
  rlstm_recurrent_kernel<<<dim3(B, 2), 4 * H, smem, s>>>(pre, 8 * H, p->w_hh_t, h0, S, H);
  DCS_LAUNCHED();
  // layer 1
  if (int e = rl_gemm_rows(h0, rows, 2 * H, p->w_ih1_t, p->bias + 8 * H, 8 * H, pre, stream)) return e;
  rlstm_recurrent_kernel<<<dim3(B, 2), 4 * H, smem, s>>>(pre, 8 * H, p->w_hh_t + (int64_t)2 * H * 4 * H, p->y, S, H);
  DCS_LAUNCHED();
  return 0;
}
