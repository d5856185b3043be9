// cconv_tc.cu — complex convolution as a real implicit GEMM on the 5th-generation tensor cores (sm_100a):
// tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM), operands staged by TMA, persistent
// warp-specialised CTAs.  This is the "bf16 mode" GEMM of every conv layer with 2*Cin % 16 == 0.
//
// Replaces apply_complex(conv_r, conv_i) / apply_complex(conv_tran_r, conv_tran_i) (complexPyTorch 0.3) at
// /root/reference/c_network.py:107-112 (encoder) and 135-147 (decoder) with the preceding torch.cat +
// complex_upsample (c_network.py:214-216) folded in; bias rule, eval-mode BN, ReLU / LReLU in the epilogue.
//
// GEMM view (include/dcsnet.h):  D[pixel, n] = sum_{tap, c} A[pixel + offset(tap), c] * W[n, tap, c]
//   M tile  = 128 output pixels of ONE sub-pixel phase = NB images x TH rows x TW cols of the phase grid
//   N       = n_pad = max(16, 2*Cout) <= 256 : one UMMA N, accumulator = n_pad TMEM columns, double-buffered
//   K step  = 64 bf16 = 128 bytes of K: 64/CK TMA boxes of the channels-last activations (CK = min(64, 2*C_src)
//             channels of one tap; zero padding and image borders come from TMA out-of-bounds fill, conv stride
//             from the tensor map's elementStrides, the channel concat from switching tensor maps) + one TMA box
//             of the K-major weight matrix (SWIZZLE_128B).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue
// (TMEM -> registers -> bias/activation -> bf16 -> global, plus optional per-(image, channel) pooling sums).
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dcs {

int validate_conv(const dcs_cconv_params* p, const char* who);

constexpr int kTileM = 128;
constexpr int kKStepBytes = 128;         // bytes of K per pipeline stage (= one 128-byte swizzle row): 64 bf16 or 32 tf32
constexpr int kTcThreads = 448;    // warp 0 TMA, warp 1 MMA, warps 2..5 A-gather (cp.async), warps 6..13 epilogue
constexpr int kGatherThreads = 128;
constexpr int kMaxStages = 8;
constexpr int kDxMaxTaps = 5;            // dx-reuse staging: up to 5 horizontal taps (k5)
constexpr uint32_t kDxABytes = 17 * 1024; // (128 + 4) rows x 128 B, rounded to the 1024-byte swizzle atom

struct TcArgs {
  int PH, PW;                 // phase grid
  int TW, TH, NB;             // tile shape, TW*TH*NB = 128
  int tiles_w, tiles_h, tiles_b, tiles_per_phase, n_tiles;
  int phases, ntaps, up_h, up_w, stride_h, stride_w;
  int C2, C2_src0, CK, ksteps, n_stages;
  int esz;                    // operand element size: 2 = fp16 / bf16 (kind::f16), 4 = fp32 read as tf32 (kind::tf32)
  int f16;                    // 16-bit operands / outputs are IEEE half (1) or bf16 (0)
  int kb;                     // 128-byte K blocks per pipeline stage: 1, or 2 (TMA mode, n_pad <= 128: one tcgen05.mma issue costs
                              // ~55 cycles and a stage's wait / fence / descriptor / commit overhead ~300, so 8 MMAs per stage
                              // instead of 4 lift the issue-bound N <= 128 layers)
  int dxm;                    // > 0: "dx-reuse" staging (TMA mode, tile = 128 consecutive pixels of ONE image row, stride_w = 1): a stage holds
                              //      ONE A box of 128 + dxm - 1 pixels x 128 bytes of K per (tap row dy, 64-channel block) and the dxm
                              //      horizontal taps read it through UMMA descriptors whose start is shifted by whole 128-byte rows (the
                              //      swizzle is a function of the absolute shared-memory address: tools/umma_shift_test), plus dxm weight
                              //      tiles.  The im2col form re-stages the A tile once per tap — dxm x the TMA rows, and the TMA row rate
                              //      (~2 cycles per 128-byte row) is what bounds these layers.
  int nrows;                  // dx-reuse: tap rows (dy values) per phase; ntaps = nrows * dxm
  int gather;                 // 1: A tiles are gathered by 4 warps with 16-byte cp.async (rows shorter than 128 B or
                              //    small N, where the per-row cost of TMA boxes dominates); 0: A tiles come from TMA boxes
  int tw_log2, th_log2;       // tile extents are powers of two
  int in_h, in_w;
  const void* src0; const void* src1;
  int n_pad, n_real, act, out_f32;
  int batch, out_h, out_w;
  int8_t dy[DCS_MAX_TAPS], dx[DCS_MAX_TAPS];
  const float* bias;
  int bias_stride;            // floats between the bias vectors of consecutive phases (0: one vector for all phases, staged in shared memory)
  void* dst;
  long long* pool;            // fixed-point pooled sums (common.cuh: pool_add) or encoded maxima (pool_max)
  int pool_max;
  unsigned long long* dbg;    // optional per-CTA wait-cycle counters (dcs_tc_set_debug_buffer), 8 words per CTA
};

constexpr int kEpiStagePayload = 2560, kEpiStageBytes = kEpiStagePayload + 32 * 8;  // per epilogue warp: transpose tile + row table
constexpr int kMaxAcc = 4;  // accumulator stages in TMEM: 4 when 4*n_pad <= 256 columns, else 2

// Division-free walk over this CTA's tiles: tile = ((phase*tiles_b + bt)*tiles_h + ht)*tiles_w + wt, advanced by a
// fixed step kept as mixed-radix digits (the per-tile divisions were a visible cost in the latency-exposed roles).
struct TileIter {
  int ph, bt, ht, wt, tile;
  int sp, sb, sh, sw, step;
  __device__ __forceinline__ void init(const TcArgs& a, int first, int stride) {
    tile = first; step = stride;
    int r = first;
    wt = r % a.tiles_w; r /= a.tiles_w; ht = r % a.tiles_h; r /= a.tiles_h; bt = r % a.tiles_b; ph = r / a.tiles_b;
    r = stride;
    sw = r % a.tiles_w; r /= a.tiles_w; sh = r % a.tiles_h; r /= a.tiles_h; sb = r % a.tiles_b; sp = r / a.tiles_b;
  }
  __device__ __forceinline__ void next(const TcArgs& a) {
    tile += step;
    wt += sw; if (wt >= a.tiles_w) { wt -= a.tiles_w; ++ht; }
    ht += sh; if (ht >= a.tiles_h) { ht -= a.tiles_h; ++bt; }
    bt += sb; if (bt >= a.tiles_b) { bt -= a.tiles_b; ++ph; }
    ph += sp;
  }
};

struct __align__(16) TcBarriers {
  float bias[256];            // first: 16-byte aligned for the epilogue's vector loads
  uint64_t full[kMaxStages], empty[kMaxStages], acc_full[kMaxAcc], acc_empty[kMaxAcc];
  uint32_t tmem_base;
};

// CTA2: the CTAs of a 2-CTA cluster work on two neighbouring M tiles of the same phase with ONE tcgen05.mma.cta_group::2
// (M = 256) per k-slice: each CTA stages its own A tile and only HALF of the weight rows, so the L2 -> shared-memory
// operand traffic per tile drops from A + B to A + B / 2 (what bounds these layers: ~45 B / clk / SM).
template <bool CTA2>
__global__ void __launch_bounds__(kTcThreads, 1)
cconv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [stage][A 16 KB | B b_rows*128 B] (1024-aligned), then barriers
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const uint32_t b_rows = CTA2 ? (uint32_t)a.n_pad / 2u : (uint32_t)a.n_pad;   // weight rows staged by this CTA
  const uint32_t b_tile = b_rows * kKStepBytes;                                  // one weight tile (b_rows x 128 bytes of K)
  const uint32_t a_bytes = a.dxm ? kDxABytes : kTileM * kKStepBytes * (uint32_t)a.kb;
  const uint32_t b_bytes = a.dxm ? b_tile * (uint32_t)a.dxm : b_tile * (uint32_t)a.kb;
  const uint32_t stage_bytes = a_bytes + b_bytes;  // multiple of 1024 (n_pad % 16 == 0 -> b_bytes % 2048 == 0)
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(base + (size_t)a.n_stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // small N: 4 accumulator stages and the two epilogue warp sets alternate TILES; large N: 2 stages, the sets split COLUMNS
  const bool tile_split = a.n_pad <= 64;
  const uint32_t n_acc = tile_split ? 4u : 2u;
  uint32_t tmem_cols = 32;
  while (tmem_cols < n_acc * (uint32_t)a.n_pad) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    const uint32_t full_count = a.gather ? 1u + kGatherThreads : 1u;  // B-TMA expect_tx arrive (+ one per gather thread)
    for (int s = 0; s < a.n_stages; ++s) { mbar_init(smem_u32(&bars->full[s]), full_count); mbar_init(smem_u32(&bars->empty[s]), 1); }
    const uint32_t n_epi_warps = tile_split ? 4u : ((a.n_pad % 32) == 0 ? 8u : 4u);
    // CTA2: the leader's acc_empty collects the epilogue warps of BOTH CTAs (its MMAs write both accumulators)
    for (int i = 0; i < kMaxAcc; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), CTA2 ? 2u * n_epi_warps : n_epi_warps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it
    if constexpr (CTA2) {   // the same warp of both CTAs, same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = threadIdx.x; i < a.n_pad; i += kTcThreads) bars->bias[i] = a.bias[i];
  const float* bias_s = bars->bias;
  // per-phase bias vectors (the merged LSTM input projections): all of them, behind the epilogue staging tiles
  float* bias_ph = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars + 1) + 8 * kEpiStageBytes);
  if (a.bias_stride)
    for (int i = threadIdx.x; i < a.phases * a.n_pad; i += kTcThreads) bias_ph[i] = a.bias[(i / a.n_pad) * a.bias_stride + (i % a.n_pad)];
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // the peer's barriers must exist before any remote arrive / commit
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int kstep_elems = kKStepBytes * a.kb / a.esz;
  const int sub_per_step = kstep_elems / a.CK;

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp loops, one lane issues)
    {
      uint32_t stage = 0, phase_bit = 0;
      const uint32_t smem_base = smem_u32(base), bar_full0 = smem_u32(&bars->full[0]), bar_empty0 = smem_u32(&bars->empty[0]);
      const uint32_t sub_bytes = (uint32_t)(kTileM * a.CK * a.esz);
      unsigned long long w_empty = 0;
      unsigned long long* dw = a.dbg ? &w_empty : nullptr;
      const long long t_start = clock64();
      TileIter it;
      it.init(a, blockIdx.x, gridDim.x);
      for (; it.tile < a.n_tiles; it.next(a)) {
        const int ph_idx = it.ph;
        const int b0 = it.bt * a.NB, j0 = it.ht * a.TH, i0 = it.wt * a.TW;
        // Single-thread control loop: keep it free of divisions / 64-bit math (every instruction is latency-exposed).
        const int8_t* dyp = a.dy + ph_idx * a.ntaps;
        const int8_t* dxp = a.dx + ph_idx * a.ntaps;
        const int ybase = j0 * a.stride_h, xbase = i0 * a.stride_w;
        const int wrow = ph_idx * a.n_pad;
        if (a.dxm) {
          // dx-reuse: stage = (tap row r, 64-channel block): one A box of 128 + dxm - 1 pixels starting at the leftmost tap, dxm weight tiles
          const uint32_t tx_bytes = (uint32_t)((kTileM + a.dxm - 1) * 128) + b_bytes;   // (TMA counts the whole box, zero fill included)
          const int xl = xbase + dxp[0];
          for (int r = 0; r < a.nrows; ++r) {
            const int yy = ybase + dyp[r * a.dxm];
            for (int cb = 0; cb < a.C2; cb += 64) {
              mbar_wait(bar_empty0 + 8u * stage, phase_bit ^ 1, dw);
              const uint32_t full = bar_full0 + 8u * stage;
              if (elect_one()) {
                const uint32_t dst = smem_base + stage * stage_bytes;
                const int kc0 = r * a.dxm * a.C2 + cb;           // weight column of tap (r, 0), channel block cb
                if constexpr (CTA2) {
                  if (cta_rank == 0) mbar_expect_tx(full, 2u * tx_bytes);
                  if (cb < a.C2_src0) tma_load_4d_2cta(dst, &tmA0, full, cb, xl, yy, b0);
                  else tma_load_4d_2cta(dst, &tmA1, full, cb - a.C2_src0, xl, yy, b0);
                  const int wr2 = wrow + (int)(cta_rank * b_rows);
                  for (int j = 0; j < a.dxm; ++j) tma_load_2d_2cta(dst + a_bytes + (uint32_t)j * b_tile, &tmB, full, kc0 + j * a.C2, wr2);
                } else {
                  mbar_expect_tx(full, tx_bytes);
                  if (cb < a.C2_src0) tma_load_4d(dst, &tmA0, full, cb, xl, yy, b0);
                  else tma_load_4d(dst, &tmA1, full, cb - a.C2_src0, xl, yy, b0);
                  for (int j = 0; j < a.dxm; ++j) tma_load_2d(dst + a_bytes + (uint32_t)j * b_tile, &tmB, full, kc0 + j * a.C2, wrow);
                }
              }
              __syncwarp();
              if (++stage == (uint32_t)a.n_stages) { stage = 0; phase_bit ^= 1; }
            }
          }
          continue;
        }
        int tap = 0, c = 0, kcol = 0;                       // running (tap, channel) position and weight column
        int y = ybase + dyp[0], x = xbase + dxp[0];
        for (int ks = 0; ks < a.ksteps; ++ks) {
          mbar_wait(bar_empty0 + 8u * stage, phase_bit ^ 1, dw);
          const uint32_t full = bar_full0 + 8u * stage;
          const bool leader = elect_one();
          if constexpr (CTA2) { if (leader && cta_rank == 0) mbar_expect_tx(full, 2u * stage_bytes); }   // both CTAs' bytes land on the leader's barrier
          else if (leader) mbar_expect_tx(full, a.gather ? b_bytes : stage_bytes);
          uint32_t dst = smem_base + stage * stage_bytes;
          for (int g = 0; g < (a.gather ? 0 : sub_per_step); ++g) {
            if (leader) {
              if constexpr (CTA2) {
                if (c < a.C2_src0) tma_load_4d_2cta(dst, &tmA0, full, c, x, y, b0);
                else tma_load_4d_2cta(dst, &tmA1, full, c - a.C2_src0, x, y, b0);
              } else {
                if (c < a.C2_src0) tma_load_4d(dst, &tmA0, full, c, x, y, b0);
                else tma_load_4d(dst, &tmA1, full, c - a.C2_src0, x, y, b0);
              }
            }
            dst += sub_bytes;
            c += a.CK;
            if (c == a.C2) {                                 // next tap (K padding repeats the last tap: finite x 0)
              c = 0;
              if (tap + 1 < a.ntaps) { ++tap; y = ybase + dyp[tap]; x = xbase + dxp[tap]; }
            }
          }
          if (leader) {
            if constexpr (CTA2) {   // this CTA's half of the weight rows
              const int wr2 = wrow + (int)(cta_rank * b_rows);
              tma_load_2d_2cta(smem_base + stage * stage_bytes + a_bytes, &tmB, full, kcol, wr2);
              if (a.kb == 2) tma_load_2d_2cta(smem_base + stage * stage_bytes + a_bytes + b_rows * kKStepBytes, &tmB, full, kcol + kKStepBytes / a.esz, wr2);
            } else {
              tma_load_2d(smem_base + stage * stage_bytes + a_bytes, &tmB, full, kcol, wrow);
              if (a.kb == 2) tma_load_2d(smem_base + stage * stage_bytes + a_bytes + b_rows * kKStepBytes, &tmB, full, kcol + kKStepBytes / a.esz, wrow);
            }
          }
          __syncwarp();
          kcol += kstep_elems;
          if (++stage == (uint32_t)a.n_stages) { stage = 0; phase_bit ^= 1; }
        }
      }
      if (a.dbg && lane == 0) { a.dbg[blockIdx.x * 8 + 0] = w_empty; a.dbg[blockIdx.x * 8 + 1] = (unsigned long long)(clock64() - t_start); }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp loops, one lane issues)
    if (!CTA2 || cta_rank == 0) {
      // instruction descriptor: D=F32, A=B=F16 (0) / BF16 (1) or TF32 (2), both K-major, N>>3 @17, M>>4 @24
      const uint32_t fmt = a.esz == 2 ? (a.f16 ? 0u : 1u) : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.n_pad >> 3) << 17) | ((uint32_t)((CTA2 ? 2 * kTileM : kTileM) >> 4) << 24);
      const uint32_t a_row_bytes = a.gather ? 128u : (uint32_t)(a.CK * a.esz);  // gathered tiles are always 128 x 128 B, SWIZZLE_128B
      // Descriptors = constant high word + (start address >> 4) in the low word; per-MMA byte offsets inside a stage
      // are precomputed once so the issue loop is a handful of 32-bit adds per MMA (single latency-exposed thread).
      const uint64_t a_desc0 = umma_desc(0, a_row_bytes), b_desc0 = umma_desc(0, 128);
      // Per-MMA descriptor offsets (16-byte units) inside a stage are COMPILE-TIME constants of the A row length: with
      // a per-thread offset table every descriptor went through vector registers and 16 R2UR moves per k-step into the
      // uniform registers UTCHMMA reads — the single issuing thread then needed ~125 cycles per MMA, twice the tensor
      // core's floor for N <= 128 (tools/umma_rate_test: 55 / 64 cycles at N = 64 / 128).
      const uint32_t smem_base = smem_u32(base), bar_full0 = smem_u32(&bars->full[0]), bar_empty0 = smem_u32(&bars->empty[0]);
      const uint32_t bar_accf0 = smem_u32(&bars->acc_full[0]), bar_acce0 = smem_u32(&bars->acc_empty[0]);
      uint32_t stage = 0, phase_bit = 0, acc = 0, acc_phase = 0;
      unsigned long long w_full = 0, w_acc = 0;
      unsigned long long *dwf = a.dbg ? &w_full : nullptr, *dwa = a.dbg ? &w_acc : nullptr;
      const long long t_start = clock64();
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(bar_acce0 + 8u * acc, acc_phase ^ 1, dwa);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)a.n_pad;
        for (int ks = 0; ks < a.ksteps; ++ks) {
          mbar_wait(bar_full0 + 8u * stage, phase_bit, dwf);
          tc_fence_after();
          const uint32_t sa16 = ((smem_base + stage * stage_bytes) & 0x3FFFFu) >> 4;  // stage start, 16-byte units
          const uint32_t sb16 = sa16 + (a_bytes >> 4);
          const uint64_t ad0 = a_desc0 + (uint64_t)sa16, bd0 = b_desc0 + (uint64_t)sb16;
          const uint32_t acc0 = ks != 0;
          if (a.dxm) {
            if (elect_one()) {
              // horizontal tap j: the A box read from row j on (descriptor start + j x 128 B), weight tile j
#pragma unroll
              for (int j = 0; j < kDxMaxTaps; ++j) {
                if (j < a.dxm) {
                  const uint64_t aj = ad0 + (uint64_t)(j * 8), bj = bd0 + (uint64_t)((uint32_t)j * (b_tile >> 4));
                  if constexpr (CTA2) {
                    tc_mma_f16_2cta(d_tmem, aj, bj, idesc, (j == 0) ? acc0 : 1u);
                    tc_mma_f16_2cta(d_tmem, aj + 2, bj + 2, idesc, 1u);
                    tc_mma_f16_2cta(d_tmem, aj + 4, bj + 4, idesc, 1u);
                    tc_mma_f16_2cta(d_tmem, aj + 6, bj + 6, idesc, 1u);
                  } else {
                    tc_mma_bf16(d_tmem, aj, bj, idesc, (j == 0) ? acc0 : 1u);
                    tc_mma_bf16(d_tmem, aj + 2, bj + 2, idesc, 1u);
                    tc_mma_bf16(d_tmem, aj + 4, bj + 4, idesc, 1u);
                    tc_mma_bf16(d_tmem, aj + 6, bj + 6, idesc, 1u);
                  }
                }
              }
              if constexpr (CTA2) {
                tc_commit_2cta(bar_empty0 + 8u * stage);
                if (ks == a.ksteps - 1) tc_commit_2cta(bar_accf0 + 8u * acc);
              } else {
                tc_commit(bar_empty0 + 8u * stage);
                if (ks == a.ksteps - 1) tc_commit(bar_accf0 + 8u * acc);
              }
            }
          } else if (elect_one()) {
            // one MMA per 32 bytes of K (16 bf16 / 8 tf32); A sub-tile g = 128 rows x a_row_bytes, 32-byte slice j inside it
#define DCS_TC_ISSUE4(O1, O2, O3)                                                        \
  do {                                                                                   \
    if (CTA2) {                                                                          \
      tc_mma_f16_2cta(d_tmem, ad0, bd0, idesc, acc0);                                    \
      tc_mma_f16_2cta(d_tmem, ad0 + (O1), bd0 + 2, idesc, 1u);                           \
      tc_mma_f16_2cta(d_tmem, ad0 + (O2), bd0 + 4, idesc, 1u);                           \
      tc_mma_f16_2cta(d_tmem, ad0 + (O3), bd0 + 6, idesc, 1u);                           \
    } else if (a.esz == 2) {                                                             \
      tc_mma_bf16(d_tmem, ad0, bd0, idesc, acc0);                                        \
      tc_mma_bf16(d_tmem, ad0 + (O1), bd0 + 2, idesc, 1u);                               \
      tc_mma_bf16(d_tmem, ad0 + (O2), bd0 + 4, idesc, 1u);                               \
      tc_mma_bf16(d_tmem, ad0 + (O3), bd0 + 6, idesc, 1u);                               \
    } else {                                                                             \
      tc_mma_tf32(d_tmem, ad0, bd0, idesc, acc0);                                        \
      tc_mma_tf32(d_tmem, ad0 + (O1), bd0 + 2, idesc, 1u);                               \
      tc_mma_tf32(d_tmem, ad0 + (O2), bd0 + 4, idesc, 1u);                               \
      tc_mma_tf32(d_tmem, ad0 + (O3), bd0 + 6, idesc, 1u);                               \
    }                                                                                    \
  } while (0)
            if (a_row_bytes == 128u) {
              DCS_TC_ISSUE4(2, 4, 6);
              if (a.kb == 2) {   // second 128-byte K block of the stage: next A sub-tile (128 rows x 128 B) and B sub-tile
                const uint64_t ad1 = ad0 + (uint64_t)((kTileM * 128) >> 4), bd1 = bd0 + (uint64_t)((b_rows * 128u) >> 4);
                if constexpr (CTA2) {
                  tc_mma_f16_2cta(d_tmem, ad1, bd1, idesc, 1u);
                  tc_mma_f16_2cta(d_tmem, ad1 + 2, bd1 + 2, idesc, 1u);
                  tc_mma_f16_2cta(d_tmem, ad1 + 4, bd1 + 4, idesc, 1u);
                  tc_mma_f16_2cta(d_tmem, ad1 + 6, bd1 + 6, idesc, 1u);
                } else {
                  tc_mma_bf16(d_tmem, ad1, bd1, idesc, 1u);
                  tc_mma_bf16(d_tmem, ad1 + 2, bd1 + 2, idesc, 1u);
                  tc_mma_bf16(d_tmem, ad1 + 4, bd1 + 4, idesc, 1u);
                  tc_mma_bf16(d_tmem, ad1 + 6, bd1 + 6, idesc, 1u);
                }
              }
            }
            else if (a_row_bytes == 64u) DCS_TC_ISSUE4(2, (kTileM * 64) >> 4, ((kTileM * 64) >> 4) + 2);
            else DCS_TC_ISSUE4((kTileM * 32) >> 4, (2 * kTileM * 32) >> 4, (3 * kTileM * 32) >> 4);
#undef DCS_TC_ISSUE4
            if constexpr (CTA2) {                   // the same barriers of BOTH CTAs (multicast commit)
              tc_commit_2cta(bar_empty0 + 8u * stage);
              if (ks == a.ksteps - 1) tc_commit_2cta(bar_accf0 + 8u * acc);
            } else {
              tc_commit(bar_empty0 + 8u * stage);     // frees the smem stage once these MMAs have read it
              if (ks == a.ksteps - 1) tc_commit(bar_accf0 + 8u * acc);  // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)a.n_stages) { stage = 0; phase_bit ^= 1; }
        }
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
      if (a.dbg && lane == 0) { a.dbg[blockIdx.x * 8 + 2] = w_full; a.dbg[blockIdx.x * 8 + 3] = w_acc; a.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_start); }
    }
  } else if (warp < 6) {
    // ===================================================================== A gather producers (4 warps, cp.async)
    // Thread g owns 16-byte chunk (g & 7) of tile rows (g >> 3) + 16*i, i = 0..7: eight lanes cover the 128 bytes of
    // K of one pixel (full 32-byte sectors), and the chunk -> (tap, channel) mapping is uniform per thread.
    if (a.gather) {
      const int g = threadIdx.x - 64;
      const int chunk = g & 7, r8 = g >> 3;
      const int epc = 16 / a.esz;                                   // elements per 16-byte chunk
      const uint32_t smem_base = smem_u32(base), bar_full0 = smem_u32(&bars->full[0]), bar_empty0 = smem_u32(&bars->empty[0]);
      const uint32_t dst_thread = (uint32_t)(r8 * 128 + ((chunk ^ (r8 & 7)) << 4));  // SWIZZLE_128B: chunk ^= row & 7
      const char* s0 = reinterpret_cast<const char*>(a.src0);
      const char* s1 = reinterpret_cast<const char*>(a.src1);
      const int C2s1 = a.C2 - a.C2_src0;
      uint32_t stage = 0, phase_bit = 0;
      TileIter it;
      it.init(a, blockIdx.x, gridDim.x);
      for (; it.tile < a.n_tiles; it.next(a)) {
        const int ph_idx = it.ph, bt = it.bt, ht = it.ht, wt = it.wt;
        int pix0[8];                                                 // source pixel index of tap (0,0) per owned row
        int ys[8], xs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = r8 + 16 * i;
          const int cc = m & (a.TW - 1), rr = (m >> a.tw_log2) & (a.TH - 1), nb = m >> (a.tw_log2 + a.th_log2);
          const int b = bt * a.NB + nb, j = ht * a.TH + rr, ii = wt * a.TW + cc;
          const bool ok = b < a.batch && j < a.PH && ii < a.PW;
          ys[i] = ok ? j * a.stride_h : -(1 << 20);                  // invalid rows fail every bounds check -> zero fill
          xs[i] = ii * a.stride_w;
          pix0[i] = (b * a.in_h + j * a.stride_h) * a.in_w + ii * a.stride_w;
        }
        const int8_t* dyp = a.dy + ph_idx * a.ntaps;
        const int8_t* dxp = a.dx + ph_idx * a.ntaps;
        int c = chunk * epc, tap = 0;                                // this thread's (tap, channel) position in K
        while (c >= a.C2) { c -= a.C2; ++tap; }
        for (int ks = 0; ks < a.ksteps; ++ks) {
          mbar_wait(bar_empty0 + 8u * stage, phase_bit ^ 1);
          const uint32_t dst0 = smem_base + stage * stage_bytes + dst_thread;
          const bool tap_ok = tap < a.ntaps;                         // K padding beyond the last tap: zeros
          const int dy = tap_ok ? dyp[tap] : 0, dx = tap_ok ? dxp[tap] : 0;
          const bool from0 = c < a.C2_src0;
          const char* sp = from0 ? s0 : s1;
          const int cs = from0 ? a.C2_src0 : C2s1, co = from0 ? c : c - a.C2_src0;
          const int dpix = dy * a.in_w + dx;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int y = ys[i] + dy, x = xs[i] + dx;
            const bool ok = tap_ok && (unsigned)y < (unsigned)a.in_h && (unsigned)x < (unsigned)a.in_w;
            const int64_t off = ok ? ((int64_t)(pix0[i] + dpix) * cs + co) * a.esz : 0;
            cp_async16(dst0 + (uint32_t)(i * 16 * 128), sp + off, ok ? 16u : 0u);
          }
          cp_async_arrive_noinc(bar_full0 + 8u * stage);
          c += kstep_elems;
          while (c >= a.C2) { c -= a.C2; ++tap; }
          if (++stage == (uint32_t)a.n_stages) { stage = 0; phase_bit ^= 1; }
        }
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps)
    // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant split the accumulator columns.
    const int ew = warp - 6;
    const int quad = warp & 3;
    const int set = ew >> 2;                          // two sets of four warps (one warp per TMEM lane quadrant each)
    const bool split = !tile_split && (a.n_pad % 32) == 0;  // large N: the sets split the accumulator columns
    const int ncols = split ? a.n_pad / 2 : a.n_pad;
    const int col0 = split ? set * ncols : 0;
    const bool active = tile_split || split || set == 0;
    // Coalescing stage of this warp (after the barriers): the TMEM layout gives a lane one tile ROW (= one pixel, whose
    // channels are contiguous in HBM), so direct stores put the 32 lanes of every store instruction into 32 different
    // lines, 16 bytes each — 8192 partial-sector requests per 128 x 256 fp32 tile, ~2 cycles each (measured with the
    // role counters: 17 k cycles of epilogue per tile against 2 k cycles of MMA for the LSTM projections).  Instead a
    // warp transposes through 2.5 KB of shared memory so that a store instruction writes whole 64 / 128-byte row pieces.
    unsigned char* est = reinterpret_cast<unsigned char*>(bars + 1) + ew * kEpiStageBytes;
    const uint32_t est_u32 = smem_u32(est);
    long long* pixs = reinterpret_cast<long long*>(est + kEpiStagePayload);
    const int m = quad * 32 + lane;                  // tile row = TMEM lane
    const int cc = m & (a.TW - 1), rr = (m >> a.tw_log2) & (a.TH - 1), nb = m >> (a.tw_log2 + a.th_log2);
    const bool warp_uniform_image = ((a.TH * a.TW) % 32) == 0;
    // tile-split: set s takes this CTA's tiles s, s+2, s+4, ... (accumulator stage = local tile index % 4)
    const int t_first = tile_split ? set : 0, t_step = tile_split ? 2 : 1;
    uint32_t acc = (uint32_t)t_first, acc_phase = 0;
    unsigned long long w_epi = 0;
    unsigned long long* dwe = (a.dbg && ew == 0 && lane == 0) ? &w_epi : nullptr;
    const long long t_start = clock64();
    int n_my_tiles = 0;
    TileIter it;
    it.init(a, blockIdx.x + t_first * gridDim.x, t_step * gridDim.x);
    for (; it.tile < a.n_tiles; it.next(a)) {
      ++n_my_tiles;
      if (active) {
        const int ph_idx = it.ph;
        const int b = it.bt * a.NB + nb, j = it.ht * a.TH + rr, i = it.wt * a.TW + cc;
        const bool valid = b < a.batch && j < a.PH && i < a.PW;
        const int ph_h = ph_idx / a.up_w, ph_w = ph_idx - ph_h * a.up_w;
        const int oy = j * a.up_h + ph_h, ox = i * a.up_w + ph_w;
        const int64_t pix = ((int64_t)b * a.out_h + oy) * a.out_w + ox;
        // output pixel of every row of this warp's quadrant (-1: outside the tensor), for the lanes that store the row
        __syncwarp();
        pixs[lane] = valid ? (long long)pix : -1ll;   // (pix is computed for every lane; `valid` masks rows outside the tensor)
        __syncwarp();
        // element offset of the rows this lane STORES: fp32 rows (lane >> 3) + 4 j (two 16-row passes, j < 8),
        // bf16 rows (lane >> 2) + 8 j (j < 4)
        int prow[8];                                   // output PIXEL index (< 2^31) of the stored rows, -1 outside the tensor
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = a.out_f32 ? (lane >> 3) + 4 * j : (lane >> 2) + 8 * (j & 3);
          const long long pv = pixs[r];
          prow[j] = pv < 0 ? -1 : (int)pv;
        }
        mbar_wait(smem_u32(&bars->acc_full[acc]), acc_phase, dwe);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)a.n_pad + (uint32_t)col0;
        for (int c = 0; c < ncols; c += 32) {
          const int nblk = min(32, ncols - c);        // 32, or a trailing 16
          const int n0 = col0 + c;
          uint32_t rg0[16], rg1[16];
          tc_ld16(taddr + (uint32_t)c, rg0);
          if (nblk == 32) tc_ld16(taddr + (uint32_t)c + 16, rg1);
          tc_ld_wait();
          // bias (vector loads) + activation; the activation is selected once per block (the epilogue is bound by its own
          // instruction stream: a per-value runtime switch cost ~8 instructions per accumulator element)
          float v[32];
          {
            const float4* b4 = reinterpret_cast<const float4*>((a.bias_stride ? bias_ph + ph_idx * a.n_pad : bias_s) + n0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bb = b4[q];
              v[4 * q] = __uint_as_float(rg0[4 * q]) + bb.x; v[4 * q + 1] = __uint_as_float(rg0[4 * q + 1]) + bb.y;
              v[4 * q + 2] = __uint_as_float(rg0[4 * q + 2]) + bb.z; v[4 * q + 3] = __uint_as_float(rg0[4 * q + 3]) + bb.w;
            }
            if (nblk == 32) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 bb = b4[4 + q];
                v[16 + 4 * q] = __uint_as_float(rg1[4 * q]) + bb.x; v[16 + 4 * q + 1] = __uint_as_float(rg1[4 * q + 1]) + bb.y;
                v[16 + 4 * q + 2] = __uint_as_float(rg1[4 * q + 2]) + bb.z; v[16 + 4 * q + 3] = __uint_as_float(rg1[4 * q + 3]) + bb.w;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) v[16 + q] = 0.f;
            }
            if (a.act == DCS_ACT_RELU) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
            } else if (a.act == DCS_ACT_LRELU) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.01f * v[q]);     // = v > 0 ? v : 0.01 v
            } else if (a.act == DCS_ACT_SIGMOID) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = sigmoidf_(v[q]);
            }
          }
          if (nblk == 32 && n0 + 32 <= a.n_real) {
            if (a.out_f32) {
              // two passes of 16 rows x 128 bytes (pitch 144): a store instruction writes 4 rows x one full 128-byte line
              float* ob = reinterpret_cast<float*>(a.dst) + n0 + (lane & 7) * 4;
#pragma unroll
              for (int pass = 0; pass < 2; ++pass) {
                if ((lane >> 4) == pass) {
                  const uint32_t wr = est_u32 + (uint32_t)(lane & 15) * 144u;
#pragma unroll
                  for (int q = 0; q < 8; ++q)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wr + q * 16), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float4 t;
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                               : "r"(est_u32 + (uint32_t)((lane >> 3) + 4 * i) * 144u + (uint32_t)(lane & 7) * 16u));
                  const int pr = prow[4 * pass + i];
                  if (pr >= 0) *reinterpret_cast<float4*>(ob + (int64_t)pr * a.n_real) = t;
                }
                __syncwarp();
              }
            } else {
              // 32 rows x 64 bytes (pitch 80): a store instruction writes 8 rows x 64 contiguous bytes
              uint32_t pk[16];
#pragma unroll
              if (a.f16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) pk[q] = pack_f16x2(v[2 * q], v[2 * q + 1]);
              } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
              }
              const uint32_t wr = est_u32 + (uint32_t)lane * 80u;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(wr + q * 16), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
              __syncwarp();
              __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(a.dst) + n0 + (lane & 3) * 8;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 t;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                             : "r"(est_u32 + (uint32_t)((lane >> 2) + 8 * i) * 80u + (uint32_t)(lane & 3) * 16u));
                if (prow[i] >= 0) *reinterpret_cast<uint4*>(ob + (int64_t)prow[i] * a.n_real) = t;
              }
              __syncwarp();
            }
          } else if (valid) {
            if (a.out_f32) {
              float* o = reinterpret_cast<float*>(a.dst) + pix * a.n_real + n0;
              if (n0 + nblk <= a.n_real) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                  if (q < nblk) *reinterpret_cast<float4*>(o + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) if (q < nblk && n0 + q < a.n_real) o[q] = v[q];
              }
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.dst) + pix * a.n_real + n0;
              if (n0 + nblk <= a.n_real) {
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) pk[q] = pack_h2_rt(v[2 * q], v[2 * q + 1], a.f16 != 0);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (8 * q < nblk) *reinterpret_cast<uint4*>(o + 8 * q) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q)
                  if (q < nblk && n0 + q < a.n_real) reinterpret_cast<unsigned short*>(o)[q] = (unsigned short)(pack_h2_rt(v[q], 0.f, a.f16 != 0) & 0xffffu);
              }
            }
          }
          if (a.pool && a.pool_max) {   // per-(image, channel) maxima: the real path's AdaptiveMaxPool2d(1) (r_network.py:11, 23)
            if (warp_uniform_image) {
              if (!valid) {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = -INFINITY;
              }
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int q = 0; q < off; ++q) {
                  const float send = up ? v[q] : v[q + off];
                  const float keep = up ? v[q + off] : v[q];
                  v[q] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
                }
              }
              const int bw = __shfl_sync(0xffffffffu, b, 0);
              if (lane < nblk && bw < a.batch && n0 + lane < a.n_real) pool_max(a.pool + (int64_t)bw * a.n_real + n0 + lane, v[0]);
            } else if (valid) {
#pragma unroll
              for (int q = 0; q < 32; ++q) if (q < nblk && n0 + q < a.n_real) pool_max(a.pool + (int64_t)b * a.n_real + n0 + q, v[q]);
            }
          } else if (a.pool) {  // numerator of the ComplexAdaptiveAvgPool2d(1) that follows (c_network.py:219)
            if (warp_uniform_image) {
              // recursive-halving transpose-reduce: 31 shuffles turn 32 columns x 32 lanes into one column sum per lane
              if (!valid) {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0.f;
              }
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int q = 0; q < off; ++q) {
                  const float send = up ? v[q] : v[q + off];
                  const float keep = up ? v[q + off] : v[q];
                  v[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              const int bw = __shfl_sync(0xffffffffu, b, 0);
              if (lane < nblk && bw < a.batch && n0 + lane < a.n_real) pool_add(a.pool + (int64_t)bw * a.n_real + n0 + lane, v[0]);
            } else if (valid) {
#pragma unroll
              for (int q = 0; q < 32; ++q) if (q < nblk && n0 + q < a.n_real) pool_add(a.pool + (int64_t)b * a.n_real + n0 + q, v[q]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (CTA2) mbar_arrive_leader(smem_u32(&bars->acc_empty[acc])); else mbar_arrive(smem_u32(&bars->acc_empty[acc])); }
      }
      acc += (uint32_t)t_step;
      if (acc >= n_acc) { acc -= n_acc; acc_phase ^= 1; }
    }
    if (dwe) { a.dbg[blockIdx.x * 8 + 5] = w_epi; a.dbg[blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - t_start); a.dbg[blockIdx.x * 8 + 7] = (unsigned long long)n_my_tiles; }
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // (pair: the leader's MMAs read the peer's shared memory and TMEM)
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// activations: channels-last complex bf16 (B, H, W, C2) -> 4-D map, box (CK, TW*sx, TH*sy, NB), element strides (1,sx,sy,1)
static CUtensorMapDataType map_dtype(int esz, int f16) {
  return esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

static int make_act_map(CUtensorMap* m, const void* ptr, int esz, int f16, int C2, int W, int H, int B, int CK, int TW, int TH, int NB, int sx, int sy, int box_w = 0) {
  EncodeTiledFn fn = encode_fn();
  DCS_REQUIRE(fn, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[4] = {(cuuint64_t)C2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C2 * esz, (cuuint64_t)W * C2 * esz, (cuuint64_t)H * W * C2 * esz};
  cuuint32_t box[4] = {(cuuint32_t)CK, (cuuint32_t)(box_w ? box_w : TW * sx), (cuuint32_t)(TH * sy), (cuuint32_t)NB};
  cuuint32_t es[4] = {1, (cuuint32_t)sx, (cuuint32_t)sy, 1};
  DCS_REQUIRE(box[1] <= 256 && box[2] <= 256, "TMA box too large (%u x %u)", box[1], box[2]);
  CUresult r = fn(m, map_dtype(esz, f16), 4, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(CK * esz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activations) failed with CUresult %d (C2=%d W=%d H=%d B=%d box=%d,%d,%d,%d)",
              (int)r, C2, W, H, B, CK, TW * sx, TH * sy, NB);
  return 0;
}

static int make_weight_map(CUtensorMap* m, const void* ptr, int esz, int f16, int Kpad, int rows, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  DCS_REQUIRE(fn, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * esz};
  cuuint32_t box[2] = {(cuuint32_t)(kKStepBytes / esz), (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, map_dtype(esz, f16), 2, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
  return 0;
}

// strips of the row-strip kernel (cconv_strip.cu): (B, H, W_units, row_elems) bf16 -> box (row_elems, box_units, 1, 1)
int make_act_map_generic(CUtensorMap* m, const void* ptr, int f16, int row_elems, int w_units, int h, int b, int box_units) {
  EncodeTiledFn fn = encode_fn();
  DCS_REQUIRE(fn, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  const cuuint64_t rb = (cuuint64_t)row_elems * 2;
  cuuint64_t dims[4] = {(cuuint64_t)row_elems, (cuuint64_t)w_units, (cuuint64_t)h, (cuuint64_t)b};
  cuuint64_t strides[3] = {rb, (cuuint64_t)w_units * rb, (cuuint64_t)h * w_units * rb};
  cuuint32_t box[4] = {(cuuint32_t)row_elems, (cuuint32_t)box_units, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, map_dtype(2, f16), 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_for((int)rb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(strip) failed with CUresult %d (row=%d w=%d h=%d b=%d box=%d)", (int)r, row_elems,
              w_units, h, b, box_units);
  return 0;
}

static unsigned long long* g_tc_dbg = nullptr;

static int pick_pow2_tile(int extent, int max_tile) {
  // largest power of two <= max_tile whose padding waste is within 4 % of the best achievable
  double best = 1e9;
  for (int t = max_tile; t >= 1; t >>= 1) best = std::min(best, (double)((extent + t - 1) / t * t) / extent);
  for (int t = max_tile; t >= 1; t >>= 1)
    if ((double)((extent + t - 1) / t * t) / extent <= best * 1.04) return t;
  return 1;
}

}  // namespace dcs

using namespace dcs;

extern "C" int dcs_cconv2d_tc_fwd(const dcs_cconv_params* p, void* stream) {
  if (int e = validate_conv(p, "dcs_cconv2d_tc_fwd")) return e;
  const int esz = is_h16(p->in_dtype) ? 2 : 4;  // fp16 / bf16 operands (kind::f16) or fp32 operands read as tf32
  const int f16 = p->in_dtype == DCS_F16 ? 1 : 0;
  DCS_REQUIRE(!is_h16(p->out_dtype) || !is_h16(p->in_dtype) || p->out_dtype == p->in_dtype,
              "dcs_cconv2d_tc_fwd: 16-bit input and output must be the same type");
  const int kblk_elems = kKStepBytes / esz;          // elements of one 128-byte K block
  int kstep_elems = kblk_elems;
  const int C2s0 = 2 * p->c0, C2s1 = 2 * p->c1, C2 = C2s0 + C2s1;
  const int CK = C2s0 >= kblk_elems ? kblk_elems : C2s0;
  const int N_ = 2 * p->cout;
  // A-operand path: TMA boxes cost ~2 cycles per box row regardless of the row length, so short rows (few channels per
  // tap: several boxes per K step) are gathered with 16-byte cp.async instead (measured: enc1 1.09 -> 0.78 ms; layers
  // with full 128-byte rows are faster through TMA).  DCS_TC_AMODE=tma|gather overrides.
  int gather = (CK * esz < 64) ? 1 : 0;
  (void)N_;
  if (const char* e = getenv("DCS_TC_AMODE")) { if (!strcmp(e, "tma")) gather = 0; else if (!strcmp(e, "gather")) gather = 1; }
  if (C2s0 % (16 / esz) || C2s1 % (16 / esz)) gather = 0;
  DCS_REQUIRE(gather || CK * esz == 32 || CK * esz == 64 || CK * esz == 128,
              "dcs_cconv2d_tc_fwd: 2*c0*sizeof(elem) must be 32, 64 or a multiple of 128 bytes (got %d); use dcs_cconv2d_fwd", C2s0 * esz);
  DCS_REQUIRE(gather || (C2s0 % CK == 0 && C2s1 % CK == 0), "dcs_cconv2d_tc_fwd: channel counts must be multiples of %d", CK);
  const int N = 2 * p->cout, n_pad = (N + 15) / 16 * 16;
  DCS_REQUIRE(n_pad <= 256, "dcs_cconv2d_tc_fwd: 2*cout must be <= 256 (got %d)", N);
  DCS_REQUIRE(((uintptr_t)p->src0 % 16 == 0) && ((uintptr_t)p->src1 % 16 == 0) && ((uintptr_t)p->weight % 16 == 0),
              "dcs_cconv2d_tc_fwd: pointers must be 16-byte aligned");

  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.PH = p->out_h / p->up_h;
  a.PW = p->out_w / p->up_w;
  a.TW = pick_pow2_tile(a.PW, std::min(kTileM, 256 / p->stride_w));  // TMA box dim (TW*stride) <= 256
  a.TH = pick_pow2_tile(a.PH, std::min(kTileM / a.TW, 256 / p->stride_h));
  a.NB = kTileM / (a.TW * a.TH);
  a.tiles_w = (a.PW + a.TW - 1) / a.TW;
  a.tiles_h = (a.PH + a.TH - 1) / a.TH;
  a.tiles_b = (p->batch + a.NB - 1) / a.NB;
  // CTA pairs (cta_group::2): 16-bit TMA-mode layers.  DCS_TC_2CTA=0 keeps single-CTA MMAs (A/B runs).
  static const bool no_2cta = getenv("DCS_TC_2CTA") && atoi(getenv("DCS_TC_2CTA")) == 0;
  const bool cta2 = !no_2cta && !gather && esz == 2 && n_pad % 16 == 0 && n_pad >= 32 && num_sms() >= 2;
  // a pair works on tiles (2q, 2q + 1) of one phase (same weights): make the tile count of a phase even (the extra tile
  // lies beyond the batch: its loads are zero-filled out of bounds, its rows are masked in the epilogue)
  if (cta2 && ((a.tiles_w * a.tiles_h * a.tiles_b) & 1)) ++a.tiles_b;
  a.tiles_per_phase = a.tiles_w * a.tiles_h * a.tiles_b;
  a.phases = p->up_h * p->up_w;
  a.n_tiles = a.tiles_per_phase * a.phases;
  a.ntaps = p->ntaps; a.up_h = p->up_h; a.up_w = p->up_w; a.stride_h = p->stride_h; a.stride_w = p->stride_w;
  a.C2 = C2; a.C2_src0 = C2s0; a.CK = CK; a.esz = esz; a.gather = gather;
  a.f16 = is_h16(p->in_dtype) ? f16 : (p->out_dtype == DCS_F16 ? 1 : 0);   // tf32 operands: the flag selects the output packing
  {
    static const bool no_kb2 = getenv("DCS_TC_NO_KB2") && atoi(getenv("DCS_TC_NO_KB2")) != 0;
    const int n_pad_ = (2 * p->cout + 15) / 16 * 16;
    a.kb = (!no_kb2 && !gather && esz == 2 && CK * esz == 128 && n_pad_ <= 128 && p->ntaps * C2 >= 4 * kblk_elems) ? 2 : 1;
    kstep_elems = kblk_elems * a.kb;
  }
  a.tw_log2 = __builtin_ctz(a.TW); a.th_log2 = __builtin_ctz(a.TH);
  a.in_h = p->in_h; a.in_w = p->in_w; a.src0 = p->src0; a.src1 = p->src1;
  const int K = p->ntaps * C2;
  a.ksteps = (K + kstep_elems - 1) / kstep_elems;
  a.n_pad = n_pad; a.n_real = N; a.act = p->act; a.out_f32 = p->out_dtype == DCS_F32;
  a.batch = p->batch; a.out_h = p->out_h; a.out_w = p->out_w;
  memcpy(a.dy, p->dy, sizeof(a.dy));
  memcpy(a.dx, p->dx, sizeof(a.dx));
  DCS_REQUIRE(p->bias, "dcs_cconv2d_tc_fwd: bias is required (pass zeros)");
  a.bias = p->bias; a.bias_stride = p->bias_phase_stride; a.dst = p->dst; a.pool = reinterpret_cast<long long*>(p->pool_sums); a.dbg = g_tc_dbg;
  a.pool_max = p->pool_mode == DCS_POOL_MAX ? 1 : 0;

  const int b_rows = cta2 ? n_pad / 2 : n_pad;
  size_t stage_bytes = ((size_t)kTileM * kKStepBytes + (size_t)b_rows * kKStepBytes) * a.kb;
  // dx-reuse staging (see TcArgs::dxm): tile = 128 consecutive pixels of one image row, taps = nrows x kw with consecutive dx
  {
    static const bool no_dxm = getenv("DCS_TC_DXM") && atoi(getenv("DCS_TC_DXM")) == 0;
    int kw = 0, nrows = 0;
    if (!no_dxm && !gather && esz == 2 && CK * esz == 128 && C2s0 % 64 == 0 && C2s1 % 64 == 0 && p->stride_w == 1 &&
        a.TW == kTileM && a.TH == 1 && a.NB == 1) {
      kw = 1;
      while (kw < p->ntaps && p->dy[kw] == p->dy[0] && p->dx[kw] == p->dx[0] + kw) ++kw;
      bool ok = kw >= 2 && kw <= kDxMaxTaps && p->ntaps % kw == 0;
      nrows = ok ? p->ntaps / kw : 0;
      for (int ph = 0; ok && ph < a.phases; ++ph)
        for (int r = 0; ok && r < nrows; ++r)
          for (int j = 0; ok && j < kw; ++j) {
            const int t = ph * p->ntaps + r * kw + j;
            ok = p->dy[t] == p->dy[ph * p->ntaps + r * kw] && p->dx[t] == p->dx[ph * p->ntaps] + j;
          }
      const size_t sb = (size_t)kDxABytes + (size_t)kw * b_rows * kKStepBytes;
      if (ok && (200 * 1024) / sb >= 2) {
        a.dxm = kw; a.nrows = nrows; a.kb = 1;
        a.ksteps = nrows * (C2 / 64);
        stage_bytes = sb;
      }
    }
  }
  int n_stages = (int)((200 * 1024) / stage_bytes);
  n_stages = std::max(2, std::min(n_stages, kMaxStages));
  a.n_stages = n_stages;
  const size_t smem = 1024 + n_stages * stage_bytes + sizeof(TcBarriers) + 8 * kEpiStageBytes + (a.bias_stride ? (size_t)a.phases * n_pad * 4 : 0);
  DCS_REQUIRE(smem <= 227 * 1024, "dcs_cconv2d_tc_fwd: per-phase bias vectors do not fit shared memory (%d phases)", a.phases);

  CUtensorMap tmA0, tmA1, tmB;
  // (the packed weight matrix is padded to whole 128-byte K blocks; a half-empty last double stage reads zeros out of bounds)
  if (int e = make_weight_map(&tmB, p->weight, esz, f16, (K + kblk_elems - 1) / kblk_elems * kblk_elems, a.phases * n_pad, b_rows)) return e;
  if (gather) {
    tmA0 = tmB; tmA1 = tmB;  // unused by the kernel in gather mode (kept valid for the descriptor prefetch)
  } else {
    const int box_w = a.dxm ? a.TW + a.dxm - 1 : 0;
    if (int e = make_act_map(&tmA0, p->src0, esz, f16, C2s0, p->in_w, p->in_h, p->batch, CK, a.TW, a.TH, a.NB, p->stride_w, p->stride_h, box_w)) return e;
    if (p->c1) {
      if (int e = make_act_map(&tmA1, p->src1, esz, f16, C2s1, p->in_w, p->in_h, p->batch, CK, a.TW, a.TH, a.NB, p->stride_w, p->stride_h, box_w)) return e;
    } else {
      tmA1 = tmA0;
    }
  }

  if (cta2) {
    DCS_CUDA(cudaFuncSetAttribute(cconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(std::min(a.n_tiles, num_sms()) & ~1), 1, 1);   // n_tiles is even
    cfg.blockDim = dim3(kTcThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DCS_CUDA(cudaLaunchKernelEx(&cfg, cconv_tc_kernel<true>, tmA0, tmA1, tmB, a));
    DCS_LAUNCHED();
    return 0;
  }
  DCS_CUDA(cudaFuncSetAttribute(cconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(a.n_tiles, num_sms());
  cconv_tc_kernel<false><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(tmA0, tmA1, tmB, a);
  DCS_LAUNCHED();
  return 0;
}

// Developer aid: when set to a device buffer of >= 8 * 148 uint64, every dcs_cconv2d_tc_fwd launch records per-CTA
// cycle counters {producer wait-empty, producer total, mma wait-full, mma wait-acc, mma total, epilogue wait, epilogue
// total, tiles}.  NULL (default) disables it.
extern "C" int dcs_tc_set_debug_buffer(void* dev_ptr) {
  g_tc_dbg = reinterpret_cast<unsigned long long*>(dev_ptr);
  return 0;
}
