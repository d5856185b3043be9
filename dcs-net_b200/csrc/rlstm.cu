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

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core recurrence for hidden = 128 (the tensor-core modes): the per-step product W_hh (512 x 128) x h (128 x 4
// sequences) as mma.sync.m16n8k16 fp16 (fp32 accumulate; N = 8 columns, the upper 4 are zero padding), the H = 128 twin of
// lstm.cu's lstm_recurrent4_mma16_kernel:
//   * 16 warps; warp w owns hidden units 8w .. 8w+7 as two M tiles: rows 0-7 / 8-15 = gates (i, f) resp. (g, o) of those
//     units, so a lane's accumulator fragment holds i, f, g, o of unit 8w + lane/4 for sequences 2 (lane%4), +1 — no gate
//     exchange; lanes with lane%4 >= 2 (padding columns) take over the odd sequence of lane - 2: one cell per thread;
//   * W_hh lives in registers as fp16 A fragments (64 registers), built from the fp32 matrix in the kernel; h is kept in
//     shared memory as fp16 in a K order that makes a lane's 32 B-fragment values contiguous (four LDS.128);
//   * gate pre-activations (input projection + biases, fp32) arrive through a ring of 1 KB bulk copies, 4 steps per stage.
// fp16 carries tf32's 11-bit significand and |h| < 1, so the products are as accurate as a TF32 tensor-core recurrence.
constexpr int kRH = 128, kRG = 4 * kRH;
constexpr int kRBlk = 4, kRStg = 3, kRPitch = kRBlk * kRG + 8;   // steps per ring stage, stages, floats per sequence per stage

// single-MUFU activations (MUFU.TANH: tanh.approx.f32, max. relative error 2^-11 — the precision of the fp16 h they produce;
// the ex2 / rcp forms they replace put two more dependent MUFUs per gate on the serial path of every step)
__device__ __forceinline__ float rl_tanh(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float rl_sigmoid(float v) { return fmaf(0.5f, rl_tanh(0.5f * v), 0.5f); }
__device__ __forceinline__ void rl_mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// pre: [dir][half][rows][256] fp32 (rows = b * S + t; column half * 256 + c = gate row (half * 256 + c) of this direction,
// gate order i, f, g, o), whh: [dir][4H][H] fp32, hout: [rows][2H] (this direction at dir * H), storage type TO.
template <typename TO>
__global__ void __launch_bounds__(512, 1) rlstm_recurrent_mma16_kernel(const float* __restrict__ pre, int64_t rows, const float* __restrict__ whh,
                                                                       TO* __restrict__ hout, int B, int S) {
  __shared__ __align__(16) __half hs[2][4][kRH];   // [buffer][sequence][permuted k]
  __shared__ uint64_t full_bar[kRStg];
  extern __shared__ __align__(16) float rl_pre_s[];   // [kRStg][4][kRPitch]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y, q0 = blockIdx.x * 4;
  uint32_t af[2][8][4];
  {
    const float* wm = whh + (int64_t)dir * kRG * kRH;
    const int t = lane & 3, g8 = lane >> 2;
#pragma unroll
    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int gate = 2 * tl + (r & 1);
          const float* wr = wm + (int64_t)(gate * kRH + 8 * warp + g8) * kRH + 16 * ks + 2 * t + 8 * (r >> 1);
          const __half2 hv = __floats2half2_rn(__ldg(wr), __ldg(wr + 1));
          af[tl][ks][r] = *reinterpret_cast<const uint32_t*>(&hv);
        }
  }
  const int u = 8 * warp + (lane >> 2);                                   // this thread's hidden unit
  const int l4 = lane & 3;
  const int seq = l4 < 2 ? 2 * l4 : 2 * (l4 - 2) + 1;                     // this thread's cell: (u, seq)
  const int nb = lane >> 2;                                               // B-fragment column = sequence (real if < 4)
  // logical k = u = 16 ks + kk lives at half position t * 32 + ks * 4 + j, t = (kk & 7) >> 1, j = (kk & 1) + 2 (kk >> 3):
  // a lane's B values of all eight k-steps are 32 consecutive halves (four LDS.128)
  const int hpos = (((u & 7) >> 1) * 32) + ((u >> 4) * 4) + (u & 1) + 2 * ((u >> 3) & 1);
  const bool q_ok = q0 + seq < B;
  TO* hq = hout + (int64_t)min(q0 + seq, B - 1) * S * (2 * kRH) + dir * kRH + u;
  float c_state = 0.f;
  for (int i = tid; i < 2 * 4 * kRH; i += 512) (&hs[0][0][0])[i] = __float2half_rn(0.f);

  const int n_blocks = (S + kRBlk - 1) / kRBlk;
  auto issue_block = [&](int blk) {
    const int s0 = blk * kRBlk, n = min(kRBlk, S - s0);
    const int t_lo = dir ? S - s0 - n : s0;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[blk % kRStg]);
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(n * 4 * kRG * sizeof(float))) : "memory");
    __syncwarp();
    if (lane < 8 * n) {
      const int half = lane & 1, sr = lane >> 1, sq = sr / n, r = sr - sq * n;
      const int q = min(q0 + sq, B - 1);
      const float* src = pre + ((int64_t)(dir * 2 + half) * rows + (int64_t)q * S + t_lo + r) * 256;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rl_pre_s + ((blk % kRStg) * 4 + sq) * kRPitch + r * kRG + half * 256);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(1024u), "r"(bar) : "memory");
    }
  };
  if (tid == 0) {
    for (int i = 0; i < kRStg; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full_bar[i])), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32)
    for (int blk = 0; blk < min(kRStg, n_blocks); ++blk) issue_block(blk);

  const int hstep = dir ? -2 * kRH : 2 * kRH, pstep = dir ? -kRG : kRG;
  hq += (int64_t)(dir ? S - 1 : 0) * (2 * kRH);
  const uint4* hrd = reinterpret_cast<const uint4*>(&hs[0][nb < 4 ? nb : 0][l4 * 32]);   // + cur * 4 * kRH halves
  __half* hwr = &hs[0][seq][0] + hpos;                                                    // + (cur ^ 1) * 4 * kRH
  int cur = 0, stage = 0;
  uint32_t phase = 0;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int nbk = min(kRBlk, S - blk * kRBlk);
    {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[stage]);
      uint32_t done = 0, spins = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        if (!done && ++spins > (1u << 22)) __trap();
      }
    }
    const float* pr = rl_pre_s + (stage * 4 + seq) * kRPitch + (dir ? (nbk - 1) * kRG : 0) + u;
    for (int r = 0; r < nbk; ++r) {
      uint4 hb[4];
      if (nb < 4) {
        const uint4* hp = hrd + cur * (4 * kRH / 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) hb[i] = hp[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) hb[i] = make_uint4(0, 0, 0, 0);
      }
      float pcur[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) pcur[g] = pr[g * kRH];
      float d[2][2][4];   // two independent accumulation chains per tile (even / odd k-steps)
#pragma unroll
      for (int tl = 0; tl < 2; ++tl)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) d[tl][c][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint4 hv = hb[ks >> 1];
        const uint32_t b0 = (ks & 1) ? hv.z : hv.x, b1 = (ks & 1) ? hv.w : hv.y;
        rl_mma_f16(d[0][ks & 1], af[0][ks], b0, b1);
        rl_mma_f16(d[1][ks & 1], af[1][ks], b0, b1);
      }
      float gi[2], gf[2], gg[2], go[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        gi[e] = d[0][0][e] + d[0][1][e];
        gf[e] = d[0][0][2 + e] + d[0][1][2 + e];
        gg[e] = d[1][0][e] + d[1][1][e];
        go[e] = d[1][0][2 + e] + d[1][1][2 + e];
      }
      const int src = lane & ~2;
      const float oi = __shfl_sync(0xffffffffu, gi[1], src), of = __shfl_sync(0xffffffffu, gf[1], src);
      const float og_ = __shfl_sync(0xffffffffu, gg[1], src), oo = __shfl_sync(0xffffffffu, go[1], src);
      const bool odd = l4 >= 2;
      const float ri = (odd ? oi : gi[0]) + pcur[0], rf = (odd ? of : gf[0]) + pcur[1];
      const float rg = (odd ? og_ : gg[0]) + pcur[2], ro = (odd ? oo : go[0]) + pcur[3];
      const float ig = rl_sigmoid(ri), fg = rl_sigmoid(rf), gt = rl_tanh(rg), ot = rl_sigmoid(ro);
      c_state = fg * c_state + ig * gt;
      const float h = ot * rl_tanh(c_state);
      hwr[(cur ^ 1) * (4 * kRH)] = __float2half_rn(h);
      if (q_ok) *hq = from_float<TO>(h);
      hq += hstep; pr += pstep; cur ^= 1;
      __syncthreads();
    }
    if (tid < 32 && blk + kRStg < n_blocks) issue_block(blk + kRStg);
    if (++stage == kRStg) { stage = 0; phase ^= 1; }
  }
}

}  // namespace dcs

using namespace dcs;

// tensor-core input projection (kind::f16): out[rows][256] fp32 = a[rows][K] * w[256][K]^T + bias, a / w of the same 16-bit type
static int rl_gemm_rows_tc(const void* a, int dtype, int64_t rows, int K, const void* w, const float* bias, float* out, void* stream) {
  dcs_cconv_params c;
  memset(&c, 0, sizeof(c));
  c.src0 = a; c.c0 = K / 2; c.c1 = 0;
  c.batch = 1; c.in_h = 1; c.in_w = (int)rows; c.out_h = 1; c.out_w = (int)rows; c.cout = 128;
  c.up_h = c.up_w = 1; c.stride_h = c.stride_w = 1; c.ntaps = 1;
  c.weight = w; c.bias = bias; c.act = DCS_ACT_NONE; c.dst = out; c.in_dtype = dtype; c.out_dtype = DCS_F32;
  return dcs_cconv2d_tc_fwd(&c, stream);
}

extern "C" int64_t dcs_rlstm_tc_workspace_bytes(int batch, int seq, int hidden) {
  if (batch <= 0 || seq <= 0 || hidden != kRH) return -1;
  return (int64_t)batch * seq * (2 * kRG * 4 + 2 * kRH * 2);     // pre [2 dirs][2 halves][rows][256] fp32 + layer-0 output [rows][2H] 16-bit
}

extern "C" int dcs_rlstm_tc_fwd(const dcs_rlstm_tc_params* p, void* stream) {
  DCS_REQUIRE(p && p->x && p->y && p->w_ih0 && p->w_ih1 && p->w_hh && p->bias && p->workspace, "dcs_rlstm_tc_fwd: null pointer");
  DCS_REQUIRE(is_h16(p->dtype), "dcs_rlstm_tc_fwd: dtype must be DCS_F16 or DCS_BF16 (the fp32 mode uses dcs_rlstm_fwd)");
  const int B = p->batch, S = p->seq, D = p->in_dim, H = p->hidden;
  DCS_REQUIRE(B > 0 && S > 0 && H == kRH && D > 0 && D % 32 == 0, "dcs_rlstm_tc_fwd: only nn.LSTM(D -> 128), D %% 32 == 0, is built (got %d -> %d)", D, H);
  DCS_REQUIRE(p->workspace_bytes >= dcs_rlstm_tc_workspace_bytes(B, S, H), "dcs_rlstm_tc_fwd: workspace too small");
  DCS_REQUIRE((int64_t)B * S < (1ll << 31), "dcs_rlstm_tc_fwd: batch*seq too large");
  const int64_t rows = (int64_t)B * S;
  float* pre = reinterpret_cast<float*>(p->workspace);
  unsigned char* h0 = reinterpret_cast<unsigned char*>(pre + 4 * rows * 256);
  const unsigned char* w0 = reinterpret_cast<const unsigned char*>(p->w_ih0);
  const unsigned char* w1 = reinterpret_cast<const unsigned char*>(p->w_ih1);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)kRStg * 4 * kRPitch * sizeof(float);
  dim3 grid((B + 3) / 4, 2);
  auto rec = [&](const float* whh, void* out) -> int {
    if (p->dtype == DCS_F16) {
      DCS_CUDA(cudaFuncSetAttribute(rlstm_recurrent_mma16_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rlstm_recurrent_mma16_kernel<__half><<<grid, 512, smem, s>>>(pre, rows, whh, reinterpret_cast<__half*>(out), B, S);
    } else {
      DCS_CUDA(cudaFuncSetAttribute(rlstm_recurrent_mma16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rlstm_recurrent_mma16_kernel<__nv_bfloat16><<<grid, 512, smem, s>>>(pre, rows, whh, reinterpret_cast<__nv_bfloat16*>(out), B, S);
    }
    DCS_LAUNCHED();
    return 0;
  };
  for (int sl = 0; sl < 4; ++sl)     // slice = dir * 2 + half: gate rows half * 256 .. + 255 of direction dir
    if (int e = rl_gemm_rows_tc(p->x, p->dtype, rows, D, w0 + (int64_t)sl * 256 * D * 2, p->bias + sl * 256, pre + (int64_t)sl * rows * 256, stream)) return e;
  if (int e = rec(p->w_hh, h0)) return e;
  for (int sl = 0; sl < 4; ++sl)
    if (int e = rl_gemm_rows_tc(h0, p->dtype, rows, 2 * H, w1 + (int64_t)sl * 256 * 2 * H * 2, p->bias + 2 * kRG + sl * 256, pre + (int64_t)sl * rows * 256, stream)) return e;
  return rec(p->w_hh + (int64_t)2 * kRG * kRH, p->y);
}

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
