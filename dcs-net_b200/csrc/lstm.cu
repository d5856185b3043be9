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
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"

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

// Recurrent kernel, 4 sequences per CTA, split-K: thread = (hidden unit u = tid / 4, K quarter kq = tid % 4) keeps the
// 4 gate rows x 16 columns of W_hh it needs in registers (64 floats) and computes, for each of the 4 sequences, the 4
// partial gate sums of its quarter with packed fp32x2 FMAs (even / odd k in the two halves).  A 12-shuffle transposing
// reduce over the 4 lanes of a unit then leaves lane kq with the complete i, f, g, o pre-activations of (unit u,
// sequence kq): every lane runs one cell update (c stays in a register), h goes to a double-buffered shared array —
// ONE __syncthreads per step, no gate exchange through shared memory, no idle lanes in the activation phase.
constexpr int kRec4Smem = 3 * 4 * (8 * kG + 8) * (int)sizeof(float);   // lstm_recurrent4_kernel's pre-activation ring

// kDbg (tools/lstm_probe.cu only; the product instantiates 0): 1 = no pre-activation loads, 2 = no h stores to HBM,
// 4 = no transcendental cell update, 8 = no recurrent FMAs, 16 = no shuffles.
template <int kDbg>
__global__ void __launch_bounds__(256, 1) lstm_recurrent4_kernel(const float* __restrict__ pre0, int64_t stride_lstm,
                                                                 int64_t stride_dir, int ld, const float* __restrict__ whh,
                                                                 float* __restrict__ hout, int B, int S) {
  __shared__ __align__(16) float hs[2][4][kH];
  const int tid = threadIdx.x, u = tid >> 2, kq = tid & 3;
  const int dir = blockIdx.y;
  const int q0 = blockIdx.x * 4;
  const int lstm = q0 / (2 * B);
  float2 w[4][8];                                   // [gate][k pair] = W_hh[gate*64 + u][16*kq + 2*i, +1]
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* wrow = whh + ((int64_t)(lstm * 2 + dir) * kG + g * kH + u) * kH + 16 * kq;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(wrow) + i);
      w[g][2 * i] = make_float2(t.x, t.y);
      w[g][2 * i + 1] = make_float2(t.z, t.w);
    }
  }
  // this lane's cell: (unit u, sequence kq)
  float* hq = hout + (int64_t)(q0 + kq) * S * (2 * kH) + dir * kH + u;
  float c_state = 0.f;
  for (int i = tid; i < 2 * 4 * kH; i += 256) (&hs[0][0][0])[i] = 0.f;
  // Input pre-activations: blocks of kBlk time steps x 4 sequences are pulled into a 3-stage shared-memory ring with
  // 1 KB bulk copies (one per (sequence, step) row, issued by warp 0, completion on an mbarrier).  Per-thread 4-byte
  // loads of the same data — registers or cp.async, any prefetch depth — took 1.2 us per step (tools/lstm_probe.cu:
  // 0.585 ms per layer with them, 0.13 ms without): 32-byte sector requests scattered over 512 streams.
  constexpr int kBlk = 8, kStg = 3, kSeqPitch = kBlk * kG + 8;      // +8 floats: the 4 sequences land on distinct banks
  extern __shared__ __align__(16) float pre_s[];                     // [kStg][4][kSeqPitch]
  __shared__ uint64_t full_bar[kStg];
  const int n_blocks = (S + kBlk - 1) / kBlk;
  const float* pre_cta = pre0 + lstm * stride_lstm + dir * stride_dir;
  auto issue_block = [&](int blk) {                                  // whole warp 0: lane i copies row i of the block
    const int s0 = blk * kBlk, n = min(kBlk, S - s0);
    const int t_lo = dir ? S - s0 - n : s0;
    const uint32_t bar = smem_u32(&full_bar[blk % kStg]);
    const int lane = tid & 31;
    if (lane == 0 && !(kDbg & 1)) mbar_expect_tx(bar, (uint32_t)(n * 4 * kG * sizeof(float)));
    __syncwarp();
    if (!(kDbg & 1) && lane < 4 * n) {
      const int sq = lane / n, r = lane - sq * n;
      const int pbq = (q0 + sq) % (2 * B);
      bulk_g2s(smem_u32(pre_s + ((blk % kStg) * 4 + sq) * kSeqPitch + r * kG), pre_cta + ((int64_t)pbq * S + t_lo + r) * ld,
               (uint32_t)(kG * sizeof(float)), bar);
    }
  };
  if (tid == 0) {
    for (int i = 0; i < kStg; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32)
    for (int blk = 0; blk < min(kStg, n_blocks); ++blk) issue_block(blk);
  const bool hi = (kq & 2) != 0, lo = (kq & 1) != 0;
  for (int step = 0; step < S; ++step) {
    const int t = dir ? S - 1 - step : step;
    const int cur = step & 1;
    const int blk = step / kBlk, stage = blk % kStg;
    const int s0 = blk * kBlk, nb = min(kBlk, S - s0);
    const int t_lo = dir ? S - s0 - nb : s0;
    if (step == s0 && !(kDbg & 1)) mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)((blk / kStg) & 1));
    float pcur[4];
    {
      const float* pr = pre_s + (stage * 4 + kq) * kSeqPitch + (t - t_lo) * kG + u;
#pragma unroll
      for (int g = 0; g < 4; ++g) pcur[g] = (kDbg & 1) ? 0.f : pr[g * kH];
    }
    // all 16 h vectors of this thread's K quarter first (one shared-memory latency per step, not one per use)
    float4 hreg[4][4];
#pragma unroll
    for (int sq = 0; sq < 4; ++sq) {
      const float4* hp = reinterpret_cast<const float4*>(&hs[cur][sq][16 * kq]);
#pragma unroll
      for (int i = 0; i < 4; ++i) hreg[sq][i] = hp[i];
    }
    float2 acc[4][4];                               // [gate][sequence] (even-k, odd-k) partial sums
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int sq = 0; sq < 4; ++sq) acc[g][sq] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (kDbg & 8) break;
#pragma unroll
      for (int sq = 0; sq < 4; ++sq) {
        const float4 h4 = hreg[sq][i];
        const float2 ha = make_float2(h4.x, h4.y), hb = make_float2(h4.z, h4.w);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          ffma2(acc[g][sq], w[g][2 * i], ha);
          ffma2(acc[g][sq], w[g][2 * i + 1], hb);
        }
      }
    }
    // transposing reduce over the 4 lanes of the unit: lane kq ends up with the sums of sequence kq
    float r[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float p[4];
#pragma unroll
      for (int sq = 0; sq < 4; ++sq) p[sq] = acc[g][sq].x + acc[g][sq].y;
      float qv[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float send = hi ? p[j] : p[2 + j];
        const float keep = hi ? p[2 + j] : p[j];
        qv[j] = keep + ((kDbg & 16) ? send : __shfl_xor_sync(0xffffffffu, send, 2));
      }
      const float send = lo ? qv[0] : qv[1];
      const float keep = lo ? qv[1] : qv[0];
      r[g] = keep + ((kDbg & 16) ? send : __shfl_xor_sync(0xffffffffu, send, 1)) + pcur[g];
    }
    float h;
    if (kDbg & 4) {
      c_state = r[1] * c_state + r[0] * r[2];
      h = r[3] * c_state;
    } else {
      const float ig = quick_sigmoid(r[0]);
      const float fg = quick_sigmoid(r[1]);
      const float gg = quick_tanh(r[2]);
      const float og = quick_sigmoid(r[3]);
      c_state = fg * c_state + ig * gg;
      h = og * quick_tanh(c_state);
    }
    hs[cur ^ 1][kq][u] = h;
    if (!(kDbg & 2)) hq[(int64_t)t * (2 * kH)] = h;
    __syncthreads();
    if (tid < 32 && step == s0 + nb - 1 && blk + kStg < n_blocks) issue_block(blk + kStg);   // every thread is past this block
  }
  if (kDbg & 2) hq[0] = c_state;
}

// Tensor-core recurrent kernel (bf16 / TF32 mode only): the per-step product W_hh (256 x 64) x h (64 x 4 sequences) as
// mma.sync.m16n8k8 TF32 (fp32 accumulate; N = 8 columns, the upper 4 are zero padding).  The FFMA kernel above is bound
// by the FP32 pipe (128 FFMA2 per thread per step, ~1000 cycles); here a warp issues 16 MMAs per step.
//   warp w owns hidden units 8w .. 8w+7 as two M tiles: rows 0-7 / 8-15 = gates (i, f) resp. (g, o) of those units, so
//   the accumulator fragment of lane l holds i, f, g, o of unit 8w + l/4 for sequences 2(l%4), 2(l%4)+1 — no gate
//   exchange.  Lanes with l%4 >= 2 (padding columns) take over the odd sequence of lane l-2 (4 shuffles): one cell per
//   thread.  W_hh lives in registers in A-fragment order (w_hh_frag, packed + tf32-rounded by the host); h is kept in
//   shared memory, tf32-rounded, in a K order that makes a lane's 16 B-fragment values contiguous (4 LDS.128).
// bare MUFU.EX2 / MUFU.RCP forms (no range fix-up code): sigmoid = 1 / (1 + 2^(-x log2 e)), tanh = 2 sigmoid(2x) - 1;
// abs. error ~3e-7; 2^(+big) = inf -> rcp = 0, 2^(-big) = 0 -> 1: both limits are exact
__device__ __forceinline__ float sigmoid_mufu(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float tanh_mufu(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.8853900817779268f * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return 2.f * r - 1.f;
}
// single-MUFU forms (MUFU.TANH: tanh.approx.f32, max. relative error 2^-11 — the precision of the fp16 h it produces):
// sigmoid(x) = 0.5 + 0.5 tanh(x / 2).  Two dependent MUFUs less per step on the serial path.
__device__ __forceinline__ float tanh_hw(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float sigmoid_hw(float v) { return fmaf(0.5f, tanh_hw(0.5f * v), 0.5f); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint4 a, const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 1) lstm_recurrent4_mma_kernel(const float* __restrict__ pre0, int64_t stride_lstm,
                                                                     int64_t stride_dir, int ld, const float* __restrict__ whh_frag,
                                                                     float* __restrict__ hout, int B, int S) {
  __shared__ __align__(16) float hs[2][4][kH];     // [buffer][sequence][permuted k], tf32-rounded
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const int q0 = blockIdx.x * 4;
  const int lstm = q0 / (2 * B);
  // A fragments: [lstm][dir][warp][tile][kstep][lane] float4
  uint4 af[2][8];
  {
    const uint4* wf = reinterpret_cast<const uint4*>(whh_frag) + ((int64_t)(lstm * 2 + dir) * 8 + warp) * 2 * 8 * 32 + lane;
#pragma unroll
    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) af[tl][ks] = __ldg(wf + (tl * 8 + ks) * 32);
  }
  const int u = 8 * warp + (lane >> 2);                                   // this thread's hidden unit
  const int l4 = lane & 3;
  const int seq = l4 < 2 ? 2 * l4 : 2 * (l4 - 2) + 1;                     // this thread's cell: (u, seq)
  const int nb = lane >> 2;                                               // B-fragment column = sequence (real if < 4)
  const int hpos = (u & 3) * 16 + (u >> 3) * 2 + ((u >> 2) & 1);          // where logical k = u lives in hs[.][.]
  float* hq = hout + (int64_t)(q0 + seq) * S * (2 * kH) + dir * kH + u;
  float c_state = 0.f;
  for (int i = tid; i < 2 * 4 * kH; i += 256) (&hs[0][0][0])[i] = 0.f;

  constexpr int kBlk = 8, kStg = 3, kSeqPitch = kBlk * kG + 8;
  extern __shared__ __align__(16) float pre_s[];                          // [kStg][4][kSeqPitch]
  __shared__ uint64_t full_bar[kStg];
  const int n_blocks = (S + kBlk - 1) / kBlk;
  const float* pre_cta = pre0 + lstm * stride_lstm + dir * stride_dir;
  auto issue_block = [&](int blk) {
    const int s0 = blk * kBlk, n = min(kBlk, S - s0);
    const int t_lo = dir ? S - s0 - n : s0;
    const uint32_t bar = smem_u32(&full_bar[blk % kStg]);
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(n * 4 * kG * sizeof(float)));
    __syncwarp();
    if (lane < 4 * n) {
      const int sq = lane / n, r = lane - sq * n;
      const int pbq = (q0 + sq) % (2 * B);
      bulk_g2s(smem_u32(pre_s + ((blk % kStg) * 4 + sq) * kSeqPitch + r * kG), pre_cta + ((int64_t)pbq * S + t_lo + r) * ld,
               (uint32_t)(kG * sizeof(float)), bar);
    }
  };
  if (tid == 0) {
    for (int i = 0; i < kStg; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32)
    for (int blk = 0; blk < min(kStg, n_blocks); ++blk) issue_block(blk);

  // One block of kBlk steps per ring stage; inside a block every address advances by a constant per step (no divisions,
  // no index arithmetic on the serial path: the recurrence is bound by the length of a warp's own instruction stream).
  const int hstep = dir ? -2 * kH : 2 * kH, pstep = dir ? -kG : kG;
  hq += (int64_t)(dir ? S - 1 : 0) * (2 * kH);
  const uint4* hrd = reinterpret_cast<const uint4*>(&hs[0][nb < 4 ? nb : 0][l4 * 16]);   // + cur * 4 * kH floats
  uint32_t* hwr = reinterpret_cast<uint32_t*>(&hs[0][seq][0]) + hpos;                     // + (cur ^ 1) * 4 * kH
  int cur = 0, stage = 0;
  uint32_t phase = 0;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int nbk = min(kBlk, S - blk * kBlk);
    mbar_wait(smem_u32(&full_bar[stage]), phase);
    const float* pr = pre_s + (stage * 4 + seq) * kSeqPitch + (dir ? (nbk - 1) * kG : 0) + u;
    for (int r = 0; r < nbk; ++r) {
      // B fragments: h of sequence nb, this lane's 16 k values (k = 8 ks + l4 + 4 half  <->  position l4*16 + 2 ks + half)
      uint4 hb[4];
      if (nb < 4) {
        const uint4* hp = hrd + cur * (4 * kH / 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) hb[i] = hp[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) hb[i] = make_uint4(0, 0, 0, 0);
      }
      float pcur[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) pcur[g] = pr[g * kH];
      // two independent accumulation chains per tile (even / odd k steps) to halve the dependent-MMA latency
      float d[2][2][4];
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
        mma_tf32(d[0][ks & 1], af[0][ks], b0, b1);
        mma_tf32(d[1][ks & 1], af[1][ks], b0, b1);
      }
      // fragment: [0],[1] = rows 0-7 (gate i resp. g) for sequences 2 l4, 2 l4 + 1; [2],[3] = rows 8-15 (gate f resp. o)
      float gi[2], gf[2], gg[2], go[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        gi[e] = d[0][0][e] + d[0][1][e];
        gf[e] = d[0][0][2 + e] + d[0][1][2 + e];
        gg[e] = d[1][0][e] + d[1][1][e];
        go[e] = d[1][0][2 + e] + d[1][1][2 + e];
      }
      // lanes l4 >= 2 hold padding columns: they take the odd sequence of lane l - 2
      const int src = lane & ~2;
      const float oi = __shfl_sync(0xffffffffu, gi[1], src), of = __shfl_sync(0xffffffffu, gf[1], src);
      const float og_ = __shfl_sync(0xffffffffu, gg[1], src), oo = __shfl_sync(0xffffffffu, go[1], src);
      const bool odd = l4 >= 2;
      const float ri = (odd ? oi : gi[0]) + pcur[0], rf = (odd ? of : gf[0]) + pcur[1];
      const float rg = (odd ? og_ : gg[0]) + pcur[2], ro = (odd ? oo : go[0]) + pcur[3];
      const float ig = sigmoid_mufu(ri), fg = sigmoid_mufu(rf), gt = tanh_mufu(rg), ot = sigmoid_mufu(ro);
      c_state = fg * c_state + ig * gt;
      const float h = ot * tanh_mufu(c_state);
      // tf32 operand for the next step: the tensor core ignores the low 13 mantissa bits, + half an ulp = round to nearest
      hwr[(cur ^ 1) * (4 * kH)] = __float_as_uint(h) + 0x1000u;
      *hq = h;
      hq += hstep; pr += pstep; cur ^= 1;
      __syncthreads();
    }
    if (tid < 32 && blk + kStg < n_blocks) issue_block(blk + kStg);
    if (++stage == kStg) { stage = 0; phase ^= 1; }
  }
}

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// fp16 variant of the tensor-core recurrence: W_hh and h as fp16 (11-bit significands, like tf32; |h| < 1, |W_hh| small:
// no range issue), m16n8k16 -> 8 instead of 16 MMAs per warp and step (the legacy tensor pipe was ~275 of the ~970
// cycles of a step), two LDS.128 instead of four for the h fragments, 32 instead of 64 weight registers.
// PT = storage type of the pre-activations: float, or __half (the tensor-core mode: the projection GEMMs then write and this
// kernel reads half the bytes — the projections are bound by their 262 MB fp32 output per layer; fp16 pre-activations carry
// the same 11-bit significand as the fp16 h / W_hh products they are added to).
template <typename PT, bool HWTANH>
__global__ void __launch_bounds__(256, 1) lstm_recurrent4_mma16_kernel(const PT* __restrict__ pre0, int64_t stride_lstm,
                                                                     int64_t stride_dir, int ld, const float* __restrict__ whh,
                                                                     float* __restrict__ hout, int B, int S) {
  __shared__ __align__(16) __half hs[2][4][kH];    // [buffer][sequence][permuted k], fp16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const int q0 = blockIdx.x * 4;
  const int lstm = q0 / (2 * B);
  // A fragments of m16n8k16 (fp16: the same 11-bit significand as tf32, half the MMAs), built from the fp32 W_hh
  // (4H x H, gate-major rows) of this (lstm, dir): tile tl rows 0-7 / 8-15 = gates (i, f) resp. (g, o) of units 8w..8w+7;
  // register r of k-step ks holds columns 16 ks + 2 t + 8 (r >> 1) + {0, 1} of row (lane >> 2) + 8 (r & 1).
  uint32_t af[2][4][4];
  {
    const float* wm = whh + (int64_t)(lstm * 2 + dir) * kG * kH;
    const int t = lane & 3, g8 = lane >> 2;
#pragma unroll
    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int gate = 2 * tl + (r & 1);
          const float* wr = wm + (int64_t)(gate * kH + 8 * warp + g8) * kH + 16 * ks + 2 * t + 8 * (r >> 1);
          const __half2 hv = __floats2half2_rn(__ldg(wr), __ldg(wr + 1));
          af[tl][ks][r] = *reinterpret_cast<const uint32_t*>(&hv);
        }
  }
  const int u = 8 * warp + (lane >> 2);                                   // this thread's hidden unit
  const int l4 = lane & 3;
  const int seq = l4 < 2 ? 2 * l4 : 2 * (l4 - 2) + 1;                     // this thread's cell: (u, seq)
  const int nb = lane >> 2;                                               // B-fragment column = sequence (real if < 4)
  // logical k = u = 16 ks + kk lives at half position t * 16 + ks * 4 + j with t = (kk & 7) >> 1, j = (kk & 1) + 2 (kk >> 3):
  // a lane's B values of all four k-steps are 16 consecutive halves (two LDS.128)
  const int hpos = (((u & 7) >> 1) * 16) + ((u >> 4) * 4) + (u & 1) + 2 * ((u >> 3) & 1);
  float* hq = hout + (int64_t)(q0 + seq) * S * (2 * kH) + dir * kH + u;
  float c_state = 0.f;
  for (int i = tid; i < 2 * 4 * kH; i += 256) (&hs[0][0][0])[i] = __float2half_rn(0.f);

  constexpr int kBlk = 8, kStg = 3, kSeqPitch = kBlk * kG + 8;
  extern __shared__ __align__(16) unsigned char pre_raw[];
  PT* pre_s = reinterpret_cast<PT*>(pre_raw);                             // [kStg][4][kSeqPitch]
  __shared__ uint64_t full_bar[kStg];
  const int n_blocks = (S + kBlk - 1) / kBlk;
  const PT* pre_cta = pre0 + lstm * stride_lstm + dir * stride_dir;
  auto issue_block = [&](int blk) {
    const int s0 = blk * kBlk, n = min(kBlk, S - s0);
    const int t_lo = dir ? S - s0 - n : s0;
    const uint32_t bar = smem_u32(&full_bar[blk % kStg]);
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(n * 4 * kG * sizeof(PT)));
    __syncwarp();
    if (lane < 4 * n) {
      const int sq = lane / n, r = lane - sq * n;
      const int pbq = (q0 + sq) % (2 * B);
      bulk_g2s(smem_u32(pre_s + ((blk % kStg) * 4 + sq) * kSeqPitch + r * kG), pre_cta + ((int64_t)pbq * S + t_lo + r) * ld,
               (uint32_t)(kG * sizeof(PT)), bar);
    }
  };
  if (tid == 0) {
    for (int i = 0; i < kStg; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32)
    for (int blk = 0; blk < min(kStg, n_blocks); ++blk) issue_block(blk);

  // One block of kBlk steps per ring stage; inside a block every address advances by a constant per step (no divisions,
  // no index arithmetic on the serial path: the recurrence is bound by the length of a warp's own instruction stream).
  const int hstep = dir ? -2 * kH : 2 * kH, pstep = dir ? -kG : kG;
  hq += (int64_t)(dir ? S - 1 : 0) * (2 * kH);
  const uint4* hrd = reinterpret_cast<const uint4*>(&hs[0][nb < 4 ? nb : 0][l4 * 16]);   // + cur * 4 * kH halves
  __half* hwr = &hs[0][seq][0] + hpos;                                                    // + (cur ^ 1) * 4 * kH
  int cur = 0, stage = 0;
  uint32_t phase = 0;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int nbk = min(kBlk, S - blk * kBlk);
    mbar_wait(smem_u32(&full_bar[stage]), phase);
    const PT* pr = pre_s + (stage * 4 + seq) * kSeqPitch + (dir ? (nbk - 1) * kG : 0) + u;
    for (int r = 0; r < nbk; ++r) {
      // B fragments: h of sequence nb, this lane's 16 k values (k = 8 ks + l4 + 4 half  <->  position l4*16 + 2 ks + half)
      uint4 hb[2];
      if (nb < 4) {
        const uint4* hp = hrd + cur * (4 * kH / 8);
        hb[0] = hp[0]; hb[1] = hp[1];
      } else {
        hb[0] = make_uint4(0, 0, 0, 0); hb[1] = make_uint4(0, 0, 0, 0);
      }
      float pcur[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) pcur[g] = to_float<PT>(pr[g * kH]);
      // two independent accumulation chains per tile (even / odd k steps)
      float d[2][2][4];
#pragma unroll
      for (int tl = 0; tl < 2; ++tl)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) d[tl][c][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint4 hv = hb[ks >> 1];
        const uint32_t b0 = (ks & 1) ? hv.z : hv.x, b1 = (ks & 1) ? hv.w : hv.y;
        mma_f16(d[0][ks & 1], af[0][ks], b0, b1);
        mma_f16(d[1][ks & 1], af[1][ks], b0, b1);
      }
      // fragment: [0],[1] = rows 0-7 (gate i resp. g) for sequences 2 l4, 2 l4 + 1; [2],[3] = rows 8-15 (gate f resp. o)
      // (four independent MMAs + a two-level add tree were measured slower: 0.537 -> 0.570 ms for the LSTM stage)
      float gi[2], gf[2], gg[2], go[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        gi[e] = d[0][0][e] + d[0][1][e];
        gf[e] = d[0][0][2 + e] + d[0][1][2 + e];
        gg[e] = d[1][0][e] + d[1][1][e];
        go[e] = d[1][0][2 + e] + d[1][1][2 + e];
      }
      // lanes l4 >= 2 hold padding columns: they take the odd sequence of lane l - 2
      const int src = lane & ~2;
      const float oi = __shfl_sync(0xffffffffu, gi[1], src), of = __shfl_sync(0xffffffffu, gf[1], src);
      const float og_ = __shfl_sync(0xffffffffu, gg[1], src), oo = __shfl_sync(0xffffffffu, go[1], src);
      const bool odd = l4 >= 2;
      const float ri = (odd ? oi : gi[0]) + pcur[0], rf = (odd ? of : gf[0]) + pcur[1];
      const float rg = (odd ? og_ : gg[0]) + pcur[2], ro = (odd ? oo : go[0]) + pcur[3];
      float ig, fg, gt, ot, h;
      if constexpr (HWTANH) {
        ig = sigmoid_hw(ri); fg = sigmoid_hw(rf); gt = tanh_hw(rg); ot = sigmoid_hw(ro);
        c_state = fg * c_state + ig * gt;
        h = ot * tanh_hw(c_state);
      } else {
        ig = sigmoid_mufu(ri); fg = sigmoid_mufu(rf); gt = tanh_mufu(rg); ot = sigmoid_mufu(ro);
        c_state = fg * c_state + ig * gt;
        h = ot * tanh_mufu(c_state);
      }
      hwr[(cur ^ 1) * (4 * kH)] = __float2half_rn(h);
      *hq = h;
      hq += hstep; pr += pstep; cur ^= 1;
      __syncthreads();
    }
    if (tid < 32 && blk + kStg < n_blocks) issue_block(blk + kStg);
    if (++stage == kStg) { stage = 0; phase ^= 1; }
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
                       float* hout, int B, int S, cudaStream_t s, const float* whh_frag = nullptr, bool pre16 = false) {
  dim3 grid(4 * B / nseq, 2);
  if (pre16) {   // (only requested with nseq == 4 and the fp16 recurrence)
    static const bool hwtanh = !(getenv("DCS_LSTM_HWTANH") && atoi(getenv("DCS_LSTM_HWTANH")) == 0);   // default on (=0: ex2 / rcp forms)
    if (hwtanh) {
      cudaFuncSetAttribute(lstm_recurrent4_mma16_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRec4Smem / 2);
      lstm_recurrent4_mma16_kernel<__half, true><<<grid, 256, kRec4Smem / 2, s>>>(reinterpret_cast<const __half*>(pre), stride_lstm, stride_dir, ld, whh, hout, B, S);
    } else {
      cudaFuncSetAttribute(lstm_recurrent4_mma16_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRec4Smem / 2);
      lstm_recurrent4_mma16_kernel<__half, false><<<grid, 256, kRec4Smem / 2, s>>>(reinterpret_cast<const __half*>(pre), stride_lstm, stride_dir, ld, whh, hout, B, S);
    }
    return;
  }
  static const bool tf32_rec = getenv("DCS_LSTM_TF32") && atoi(getenv("DCS_LSTM_TF32")) != 0;   // A/B switch (default: fp16 MMAs)
  if (nseq == 4 && whh_frag && !tf32_rec) {
    cudaFuncSetAttribute(lstm_recurrent4_mma16_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRec4Smem);
    lstm_recurrent4_mma16_kernel<float, false><<<grid, 256, kRec4Smem, s>>>(pre, stride_lstm, stride_dir, ld, whh, hout, B, S);
  } else if (nseq == 4 && whh_frag) {
    cudaFuncSetAttribute(lstm_recurrent4_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRec4Smem);
    lstm_recurrent4_mma_kernel<<<grid, 256, kRec4Smem, s>>>(pre, stride_lstm, stride_dir, ld, whh_frag, hout, B, S);
  } else if (nseq == 4) {
    cudaFuncSetAttribute(lstm_recurrent4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRec4Smem);
    lstm_recurrent4_kernel<0><<<grid, 256, kRec4Smem, s>>>(pre, stride_lstm, stride_dir, ld, whh, hout, B, S);
  }
  else lstm_recurrent_kernel<2><<<grid, 256, 0, s>>>(pre, stride_lstm, stride_dir, ld, whh, hout, B, S);
}

// tensor-core input projections (tf32 operands read from fp32 memory), `nsl` weight slices in ONE launch:
// out[row][sl][256] = a[row][K] * w[sl][256][K]^T + bias[sl][256].  The slices are the "phases" of the conv kernel with
// up_w = nsl (output pixel row * nsl + sl reads source pixel row; every phase has the single tap (0, 0)), so the A tile of a
// row block is re-read from L2, not HBM, and the launch count drops from 8 to 3 per ComplexLSTM.
static int gemm_rows_tc(const float* a, int64_t rows, int K, const float* w_t, const float* bias, void* out, int nsl, void* stream, bool out16 = false) {
  dcs_cconv_params c;
  memset(&c, 0, sizeof(c));
  c.src0 = a; c.c0 = K / 2; c.c1 = 0;
  c.batch = 1; c.in_h = 1; c.in_w = (int)rows; c.out_h = 1; c.out_w = (int)rows * nsl; c.cout = kG / 2;
  c.up_h = 1; c.up_w = nsl; c.stride_h = c.stride_w = 1; c.ntaps = 1;
  c.weight = w_t; c.bias = bias; c.bias_phase_stride = kG; c.act = DCS_ACT_NONE; c.dst = out; c.in_dtype = DCS_F32; c.out_dtype = out16 ? DCS_F16 : DCS_F32;
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
    DCS_REQUIRE(is_dtype(p->in_dtype), "dcs_clstm_fwd: bad in_dtype");
    if (p->in_dtype == DCS_BF16) lstm_deinterleave_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)p->x, w.xp, rows, D);
    else if (p->in_dtype == DCS_F16) lstm_deinterleave_kernel<__half><<<g, 256, 0, s>>>((const __half*)p->x, w.xp, rows, D);
    else lstm_deinterleave_kernel<float><<<g, 256, 0, s>>>((const float*)p->x, w.xp, rows, D);
    DCS_LAUNCHED();
  }
  // NSEQ must divide 2B; 2 sequences per CTA (two co-resident CTAs per SM) unless that overflows the machine
  // 4 sequences per CTA (split-K kernel) whenever a CTA's sequences share one lstm (2B % 4 == 0); else the 2-sequence kernel
  int nseq = (2 * B) % 4 == 0 ? 4 : 2;
  if (p->seqs_per_cta == 2) nseq = 2;
  const float* whh1 = p->w_hh + (int64_t)4 * kG * kH;
  const int64_t rows2 = 2 * rows;
  if (p->w_ih0_t && p->w_ih1_t) {
    // ---- tensor-core (tf32) input projections.  layer 0: pre[(part,b,s)][lstm][dir][4H], ONE launch (4 weight slices)
    // fp16 pre-activations with the fp16 recurrence (DCS_LSTM_PRE32=1 keeps fp32 pre-activations: A/B runs)
    static const bool pre32 = getenv("DCS_LSTM_PRE32") && atoi(getenv("DCS_LSTM_PRE32")) != 0;
    static const bool tf32_rec_ = getenv("DCS_LSTM_TF32") && atoi(getenv("DCS_LSTM_TF32")) != 0;
    const bool pre16 = !pre32 && !tf32_rec_ && nseq == 4 && p->w_hh_frag;
    __half* pre_h = reinterpret_cast<__half*>(w.pre);
    if (int e = gemm_rows_tc(w.xp, rows2, D, p->w_ih0_t, p->bias, w.pre, 4, stream, pre16)) return e;
    launch_rec(nseq, w.pre, 2 * kG, kG, 4 * kG, p->w_hh, w.h0, B, S, s, p->w_hh_frag, pre16);
    DCS_LAUNCHED();
    // layer 1: per lstm, rows (part,b,s) of h0[lstm] -> pre1[lstm][(part,b,s)][dir][4H], one launch per lstm (2 slices)
    for (int l = 0; l < 2; ++l)
      if (int e = gemm_rows_tc(w.h0 + (int64_t)l * rows2 * 2 * kH, rows2, 2 * kH, p->w_ih1_t + (int64_t)l * 2 * kG * 2 * kH,
                               p->bias + 4 * kG + l * 2 * kG, pre16 ? (void*)(pre_h + (int64_t)l * rows2 * 2 * kG) : (void*)(w.pre + (int64_t)l * rows2 * 2 * kG),
                               2, stream, pre16)) return e;
    launch_rec(nseq, w.pre, rows2 * 2 * kG, kG, 2 * kG, whh1, w.h1, B, S, s,
               p->w_hh_frag ? p->w_hh_frag + (int64_t)4 * kG * kH : nullptr, pre16);
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
