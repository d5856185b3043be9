// cconv_strip.cu — "row-strip" implicit-GEMM complex convolution for the few-channel layers (encoder[1], decoder[4],
// decoder[5]) on the 5th-generation tensor cores: tcgen05.mma (bf16, fp32 accumulators in TMEM), TMA, mbarriers.
//
// Replaces apply_complex(conv_r, conv_i) / apply_complex(conv_tran_r, conv_tran_i) (complexPyTorch 0.3) at
// /root/reference/c_network.py:107-112 and 135-147 together with torch.cat + complex_upsample (c_network.py:214-216),
// eval-mode ComplexBatchNorm2d and ComplexReLU / ComplexLReLU, for layers whose per-pixel K slice is short.
//
// Why a second tensor-core kernel: the general kernel (cconv_tc.cu) re-stages the A operand once per tap and per
// sub-pixel phase (an im2col in shared memory), so a layer with C2 = 16..64 moves 16..49x its input through the
// L2 -> SM path and is bound by it.  Here ONE TMA box per source row — a strip of 128 (+halo) pixels, channels-last,
// so a strip row IS a GEMM A row — serves every tap that touches that row:
//   * a horizontal tap offset dx is the SAME strip read through a UMMA descriptor whose start address is shifted by
//     dx rows (the 32/64/128-byte swizzle is a function of the absolute shared-memory address, so a row-shifted
//     descriptor reads exactly what TMA wrote — probed on hardware by tools/umma_shift_test.cu);
//   * stride 2 pairs two pixels into one strip row, the tap parity selects the 32-byte K slice inside the row;
//   * a vertical tap offset dy is another slot of a ring of source rows: walking down the image, each source row is
//     loaded once per (column strip, phase group) and reused by every output row that needs it;
//   * the weights of the whole layer stay resident in shared memory (one bulk copy per CTA).
// The host (dcs-net_b200/packing.py: StripConv) flattens the layer into a table of MMA "items"
// {ring row, A descriptor offset, B block, accumulator column, first}; the issuing warp just walks the table.
//
// Warp roles: warp 0 = TMA producer, warps 1-2 = MMA issuers (warp 1 also allocates TMEM), warps 3.. = 4 / 8 / 16 epilogue
// warps (TMEM lane quadrant = warp % 4): bias + activation + bf16 + per-(image, channel) pooling sums, or the mask tail.
#include <cuda.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dcs {

int make_act_map_generic(CUtensorMap* m, const void* ptr, int f16, int row_elems, int w_units, int h, int b, int box_units);

constexpr int kStripM = 128;
// threads = warp 0 TMA + warps 1, 2 MMA issuers + kEpi epilogue warps (4, 8 or 16: kEpi / 4 warps per TMEM lane quadrant,
// which split the accumulator's 32-column chunks; the tail epilogue is transcendental-heavy and always uses 16).
// TWO issuing warps: tools/umma_rate_test.cu measures 54.8 cycles per M = 128, N <= 64 MMA from one issuing thread but 39-48
// from two (the tensor core's own floor, 32 + N/4 cycles: A and B streamed from shared memory at 128 B/clk), so a single
// issuer — plus its ~15 instructions of descriptor arithmetic per item — was the bound of every small-N layer.
constexpr int kStripIssuers = 2;
constexpr int kStripAcc = 4;              // accumulator stages in TMEM: row tile g uses stage g % 4, issuer g % 2
constexpr int kStripEpi0 = 1 + kStripIssuers;   // first epilogue warp
constexpr int strip_threads(int epi_warps) { return 32 * (kStripEpi0 + epi_warps); }
constexpr int kStripMaxRing = 16;
constexpr int kStripMaxItems = 64;    // MMA items of one phase group; the table travels in the kernel parameters

struct StripGroup {   // one phase group = one launch
  int dy_min, n_dy, ph0, x_min;
  uint32_t w_bytes;
  const unsigned char* w_ptr;
};

struct StripArgs {
  int batch, PH, PW;
  int out_h, out_w, up_h, up_w, n_real, run_log2;
  int s_h;
  int n_strips, n_chunks, chunk_rows, n_units;
  StripGroup grp;
  int R;
  uint32_t slot_bytes, src1_off, tx_bytes, w_smem_bytes;
  int has_src1;
  uint32_t row_bytes0, row_bytes1;
  int n_mma, act;
  int f16;                 // operands / output are IEEE half (1) or bf16 (0)
  const float* bias;
  __nv_bfloat16* dst;      // 16-bit storage (either type)
  long long* pool;         // fixed-point pooled sums (common.cuh: pool_add) or encoded maxima (pool_max)
  int pool_max;
  dcs_strip_tail tail;  // kTail instances only: decoder[6] + bound_cRM x2 + mask combine epilogue
  // The item table lives in the constant bank (kernel parameters) and the issue loop is fully unrolled (kNdy ring
  // rows x kIpr items per row are template parameters), so every item field is a constant-bank operand of a uniform
  // add: ~6 instructions per MMA on the single issuing warp.  (Walking a table in shared memory cost ~40 dependent
  // instructions, ~200 cycles, per MMA and made the issuing warp the bottleneck.)
  uint4 items[kStripMaxItems];        // [drow][kIpr] {a_off16, b_off16, d_col | drow << 16 | flags << 24, -}
};

struct __align__(8) StripBarriers {
  uint64_t full[kStripMaxRing], empty[kStripMaxRing], acc_full[kStripAcc], acc_empty[kStripAcc], wbar;
  uint32_t tmem_base;
};

// recursive-halving transpose-reduce: 31 shuffles turn 32 columns x 32 lanes into one column sum per lane
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
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
  return v[0];
}

struct UnitIter {  // unit u of a phase group -> (image, column strip, row chunk)
  int b, x0, j0, j1;
  __device__ __forceinline__ void set(const StripArgs& a, int u) {
    const int chunk = u % a.n_chunks, t = u / a.n_chunks;
    const int strip = t % a.n_strips;
    b = t / a.n_strips;
    x0 = strip * kStripM;
    j0 = chunk * a.chunk_rows;
    j1 = min(a.PH, j0 + a.chunk_rows);
  }
};

// Walks this CTA's row tiles in order across its units.  `lo` = index (in the CTA's stream of ring rows) of the tile's first
// source row, `slot` = lo % R; after the last tile `lo` = total number of rows and valid = false.
struct TileCursor {
  UnitIter un;
  int u, stride, j, n_dy;
  uint32_t lo, slot;
  bool valid;
  __device__ __forceinline__ void init(const StripArgs& a, int first, int step, int ndy) {
    u = first; stride = step; n_dy = ndy; lo = 0; slot = 0;
    valid = u < a.n_units;
    if (valid) { un.set(a, u); j = un.j0; }
  }
  __device__ __forceinline__ void advance(const StripArgs& a) {
    if (!valid) return;
    uint32_t d;
    if (++j < un.j1) d = (uint32_t)a.s_h;
    else {
      d = (uint32_t)n_dy;
      u += stride;
      valid = u < a.n_units;
      if (valid) { un.set(a, u); j = un.j0; }
    }
    lo += d; slot += d;
    while (slot >= (uint32_t)a.R) slot -= (uint32_t)a.R;
  }
};

template <int kCols, int kNdy, int kIpr, int kEpi = 4, bool kTail = false>
__global__ void __launch_bounds__(strip_threads(kEpi), 1)
cconv_strip_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const StripArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [ring: R slots][weights][items][column bias][barriers]
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* w_s = base + (size_t)a.R * a.slot_bytes;
  float* bias_col = reinterpret_cast<float*>(w_s + a.w_smem_bytes);
  StripBarriers* bars = reinterpret_cast<StripBarriers*>(bias_col + kCols);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_in_grp = blockIdx.x, ctas_in_grp = gridDim.x;
  const StripGroup& G = a.grp;
  constexpr uint32_t tmem_cols = kStripAcc * kCols < 32 ? 32u : (uint32_t)(kStripAcc * kCols);  // kCols is a power of two
  static_assert(kStripAcc * kCols <= 512, "accumulator stages exceed TMEM");

  if (threadIdx.x == 0) {
    // a ring slot is free again when BOTH issuers have released its row
    for (int s = 0; s < a.R; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), kStripIssuers); }
    for (int i = 0; i < kStripAcc; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), kEpi); }
    mbar_init(smem_u32(&bars->wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = threadIdx.x; c < kCols; c += blockDim.x) bias_col[c] = kTail ? 0.f : a.bias[c & (a.n_real - 1)];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    const bool leader = elect_one();
    if (leader) {  // the layer's weights: one bulk copy per CTA (<= 32 KB pieces)
      mbar_expect_tx(smem_u32(&bars->wbar), G.w_bytes);
      for (uint32_t off = 0; off < G.w_bytes; off += 32768u)
        bulk_g2s(smem_u32(w_s) + off, G.w_ptr + off, min(32768u, G.w_bytes - off), smem_u32(&bars->wbar));
    }
    __syncwarp();
    uint32_t slot = 0, par = 0;
    const uint32_t ring = smem_u32(base), bar_full0 = smem_u32(&bars->full[0]), bar_empty0 = smem_u32(&bars->empty[0]);
    UnitIter un;
    for (int u = cta_in_grp; u < a.n_units; u += ctas_in_grp) {
      un.set(a, u);
      const int nrows = (un.j1 - 1 - un.j0) * a.s_h + kNdy;
      const int y0 = un.j0 * a.s_h + G.dy_min, x = un.x0 + G.x_min;
      for (int r = 0; r < nrows; ++r) {
        mbar_wait(bar_empty0 + 8u * slot, par ^ 1);
        if (elect_one()) {
          const uint32_t full = bar_full0 + 8u * slot, dst = ring + slot * a.slot_bytes;
          mbar_expect_tx(full, a.tx_bytes);
          tma_load_4d(dst, &tmA0, full, 0, x, y0 + r, un.b);
          if (a.has_src1) tma_load_4d(dst + a.src1_off, &tmA1, full, 0, x, y0 + r, un.b);
        }
        __syncwarp();
        if (++slot == (uint32_t)a.R) { slot = 0; par ^= 1; }
      }
    }
  } else if (warp < kStripEpi0) {
    // ===================================================================== MMA issuers (2 warps; whole warp loops, one lane issues)
    // Issuer w owns the row tiles g with g % 2 == w (accumulator stage g % 4).  Both walk the same tile sequence; each waits
    // for the source rows its tiles need and releases a ring row once ITS last tile that reads it has been issued
    // (rows below the first row of its next own tile); the slot's empty barrier counts both issuers.
    const int w_iss = warp - 1;
    const uint32_t fmt = a.f16 ? 0u : 1u;   // A / B format: F16 (0) or BF16 (1)
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.n_mma >> 3) << 17) | ((uint32_t)(kStripM >> 4) << 24);
    const uint64_t a_desc_c0 = umma_desc(0, a.row_bytes0), a_desc_c1 = umma_desc(0, a.row_bytes1), b_desc_c = umma_desc(0, 32);
    const uint32_t ring16 = (smem_u32(base) & 0x3FFFFu) >> 4, slot16 = a.slot_bytes >> 4, w16 = (smem_u32(w_s) & 0x3FFFFu) >> 4;
    const uint32_t bar_full0 = smem_u32(&bars->full[0]), bar_empty0 = smem_u32(&bars->empty[0]);
    const uint32_t b_lo = (uint32_t)b_desc_c + w16, b_hi = (uint32_t)(b_desc_c >> 32);
    const uint32_t R = (uint32_t)a.R;
    uint32_t ws = 0, wp = 0, gw = 0;   // next full barrier to wait on (slot, parity) / rows waited so far
    uint32_t fs = 0, rel = 0;          // next slot to release / rows released so far
    mbar_wait(smem_u32(&bars->wbar), 0);
    TileCursor cur, ahead;             // tile g and tile g + 2 (this issuer's next own tile)
    cur.init(a, cta_in_grp, ctas_in_grp, kNdy);
    ahead.init(a, cta_in_grp, ctas_in_grp, kNdy);
    ahead.advance(a); ahead.advance(a);
    for (uint32_t g = 0; cur.valid; cur.advance(a), ahead.advance(a), ++g) {
      if ((int)(g & 1u) != w_iss) continue;
      const uint32_t acc = g & (kStripAcc - 1), accp = (g / kStripAcc) & 1u;
      mbar_wait(smem_u32(&bars->acc_empty[acc]), accp ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)kCols;
      // Two issue schedules (measured per instance on B200): whole-tile batches win where a tile has many ring rows with
      // few items each or the ring has slack (enc0, enc1, dec4); per-row issue keeps more overlap with the loads where the
      // ring is tight (enc2: R = n_dy + stride) and for the 3-row decoders.
      constexpr bool kBatchIssue = (kNdy == 7 || kNdy == 2);
      if constexpr (kBatchIssue) {
      // Wait for ALL source rows of the tile first (in steady state they are already there), then issue the whole tile's
      // MMAs in ONE elected region: a wait / fence / elect / sync round per ring row cost ~150-300 cycles of the issuing
      // thread against ~55 per MMA (the same stage overhead that bounded cconv_tc).
      {
        const uint32_t need = cur.lo + (uint32_t)(kNdy - 1);
        if (gw <= need) {
          while (gw <= need) {
            mbar_wait(bar_full0 + 8u * ws, wp);
            if (++ws == R) { ws = 0; wp ^= 1; }
            ++gw;
          }
          tc_fence_after();
        }
      }
      uint32_t a_lo[kNdy];
#pragma unroll
      for (int dg = 0; dg < kNdy; ++dg) {
        uint32_t sl = cur.slot + (uint32_t)dg;
        if (sl >= R) sl -= R;
        a_lo[dg] = (uint32_t)a_desc_c0 + ring16 + sl * slot16;   // descriptor low word = LBO | (row base >> 4)
      }
      if (elect_one()) {
#pragma unroll
        for (int dg = 0; dg < kNdy; ++dg) {
#pragma unroll
          for (int q = 0; q < kIpr; ++q) {
            const uint4 item = a.items[dg * kIpr + q];
            const uint64_t ad = ((uint64_t)item.w << 32) | (uint64_t)(a_lo[dg] + item.x);
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + item.y);
            tc_mma_bf16(d_tmem + (item.z & 0xffffu), ad, bd, idesc, (item.z & (1u << 24)) ? 0u : 1u);
          }
        }
      }
      __syncwarp();
      } else {
      // items are sorted by ring row (drow): wait for a source row once, then issue its MMAs back to back
#pragma unroll
      for (int dg = 0; dg < kNdy; ++dg) {
        const uint32_t need = cur.lo + (uint32_t)dg;
        if (gw <= need) {
          while (gw <= need) {
            mbar_wait(bar_full0 + 8u * ws, wp);
            if (++ws == R) { ws = 0; wp ^= 1; }
            ++gw;
          }
          tc_fence_after();
        }
        uint32_t sl = cur.slot + (uint32_t)dg;
        if (sl >= R) sl -= R;
        const uint32_t a16 = ring16 + sl * slot16;
        // (per-item `desc + (a16 + item.x)` kept every descriptor in vector registers: ~5 R2UR moves per MMA on the
        // issuing thread; with per-row bases the item offsets are uniform adds of constant-bank operands)
        // descriptor = {constant high word, low word = LBO | (address >> 4)}: the row's base goes through ONE vector ->
        // uniform move, every item adds constant-bank operands (offsets, and the A high word pre-computed by the host)
        const uint32_t a_lo = (uint32_t)a_desc_c0 + a16;
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < kIpr; ++q) {
            const uint4 item = a.items[dg * kIpr + q];
            const uint64_t ad = ((uint64_t)item.w << 32) | (uint64_t)(a_lo + item.x);
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + item.y);
            tc_mma_bf16(d_tmem + (item.z & 0xffffu), ad, bd, idesc, (item.z & (1u << 24)) ? 0u : 1u);
          }
        }
        __syncwarp();
      }
      }
      const uint32_t upto = ahead.lo;                 // first row of this issuer's next own tile (or the total row count)
      // Never release a row this issuer has not waited for: a parity wait on a phase that is two completions old would
      // spin for ever (rows only the other issuer's tile reads, at unit boundaries).  They are in flight for that tile.
      while (gw < upto) {
        mbar_wait(bar_full0 + 8u * ws, wp);
        if (++ws == R) { ws = 0; wp ^= 1; }
        ++gw;
      }
      const uint32_t nfree = upto - rel;
      if (elect_one()) {
        tc_commit(smem_u32(&bars->acc_full[acc]));
        uint32_t f = fs;
        for (uint32_t i = 0; i < nfree; ++i) { tc_commit(bar_empty0 + 8u * f); if (++f == R) f = 0; }
      }
      __syncwarp();
      fs += nfree; while (fs >= R) fs -= R;
      rel = upto;
    }
  } else {
    // ===================================================================== epilogue (4 warps, one TMEM lane quadrant each)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int run_len = 1 << a.run_log2;                 // columns of one output row run = up_w * n_real
    uint32_t acc = 0, accp = 0;                          // row tile g of this CTA uses stage g % kStripAcc
    UnitIter un;
    for (int u = cta_in_grp; u < a.n_units; u += ctas_in_grp) {
      un.set(a, u);
      const int x = un.x0 + m;
      const bool valid = x < a.PW;
      if constexpr (kTail) {
        // decoder[6] (N = 2) fused with the mask tail: columns = (ph, 8 output pixels of this strip row, re/im), i.e.
        // each 16-column half is 8 consecutive complex64 values of output row 2j + ph.  raw -> bound_cRM
        // (c_network.py:225) -> bound_cRM (network_functions.py:394) -> Y (.) mask -> Y - N (396-397) or Y (.) M (434).
        static_assert(!kTail || kCols == 32, "tail epilogue expects 2 phase rows x 8 pixels x (re, im)");
        const dcs_strip_tail& tl = a.tail;
        const bool exact = tl.exact_polar != 0;
        // 16 epilogue warps: warp (quad, sub) owns phase row ph = sub / 2 and output pixels 4 (sub % 2) .. +3 of each lane
        const int tsub = (warp - kStripEpi0) >> 2, ph = tsub >> 1, half = tsub & 1;
        for (int j = un.j0; j < un.j1; ++j) {
          mbar_wait(smem_u32(&bars->acc_full[acc]), accp);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)kCols + (uint32_t)(ph * 16 + half * 8);
          uint32_t rg[8];
          tc_ld8(taddr, rg);
          tc_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
          if (valid && tl.combine >= DCS_COMBINE_DR) {
            // real path (r_network.py:172, network_functions.py:286-305 / 338-342): m = sigmoid(logit); |S| = |Y| m (dr) or
            // |Y| - |Y| m (drs); the waveform takes the NOISY phase atan2(Im Y, Re Y + eps), so S = |S| e^{j phase} is formed
            // here and the iSTFT reads it without a polar round trip
            const int64_t o = ((int64_t)un.b * a.out_h + 2 * j + ph) * a.out_w + (int64_t)x * 8 + 4 * half;   // element index
            const float4* yp = reinterpret_cast<const float4*>(tl.noisy_spec) + (o >> 1);
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const float4 y2 = __ldg(yp + e2);
              float mk[2];
              float2 cl[2], ns[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float logit = __uint_as_float(rg[4 * e2 + 2 * h]) + tl.bias_re;
                const float2 yv = h ? make_float2(y2.z, y2.w) : make_float2(y2.x, y2.y);
                const float m = exact ? sigmoidf_(logit) : fast_sigmoid(logit);
                const float mag = hypotf(yv.x, yv.y);                       // torch.abs(complex64)
                float cs, sn;
                if (exact) { const float th = atan2f(yv.y, yv.x + tl.atan2_eps); cs = cosf(th); sn = sinf(th); }
                else {
                  const float xr = yv.x + tl.atan2_eps, hy = hypotf(xr, yv.y);
                  const float inv = hy > 0.f ? 1.f / hy : 0.f;
                  cs = hy > 0.f ? xr * inv : 1.f; sn = yv.y * inv;
                }
                const float nm = mag * m;                                    // masked magnitude
                const float cm = tl.combine == DCS_COMBINE_DRS ? mag - nm : nm;
                mk[h] = m; cl[h] = make_float2(cm * cs, cm * sn); ns[h] = make_float2(nm * cs, nm * sn);
              }
              const int64_t q = (o >> 1) + e2;
              reinterpret_cast<float4*>(tl.clean_spec)[q] = make_float4(cl[0].x, cl[0].y, cl[1].x, cl[1].y);
              if (tl.noise_spec && tl.combine == DCS_COMBINE_DRS) reinterpret_cast<float4*>(tl.noise_spec)[q] = make_float4(ns[0].x, ns[0].y, ns[1].x, ns[1].y);
              if (tl.mask) reinterpret_cast<float2*>(tl.mask)[q] = make_float2(mk[0], mk[1]);
            }
          } else if (valid) {
            const int64_t o = ((int64_t)un.b * a.out_h + 2 * j + ph) * a.out_w + (int64_t)x * 8 + 4 * half;   // complex index
            const float4* yp = reinterpret_cast<const float4*>(tl.noisy_spec) + (o >> 1);
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {      // two output pixels per 16-byte access
              const float4 y2 = __ldg(yp + e2);
              float2 raw[2], m1[2], m2[2], cl[2], ns[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c = 4 * e2 + 2 * h;
                raw[h] = make_float2(__uint_as_float(rg[c]) + tl.bias_re, __uint_as_float(rg[c + 1]) + tl.bias_im);
                m1[h] = exact ? bound_crm_dev(raw[h], tl.atan2_eps, true) : bound_crm_mufu(raw[h], tl.atan2_eps);
                m2[h] = exact ? bound_crm_dev(m1[h], tl.atan2_eps, true) : bound_crm_mufu(m1[h], tl.atan2_eps);
                const float2 yv = h ? make_float2(y2.z, y2.w) : make_float2(y2.x, y2.y);
                const float2 pr = cmul(yv, m2[h]);
                ns[h] = pr;
                cl[h] = tl.combine == DCS_COMBINE_DCS ? make_float2(yv.x - pr.x, yv.y - pr.y) : pr;
              }
              const int64_t q = (o >> 1) + e2;
              reinterpret_cast<float4*>(tl.clean_spec)[q] = make_float4(cl[0].x, cl[0].y, cl[1].x, cl[1].y);
              if (tl.noise_spec && tl.combine == DCS_COMBINE_DCS) reinterpret_cast<float4*>(tl.noise_spec)[q] = make_float4(ns[0].x, ns[0].y, ns[1].x, ns[1].y);
              if (tl.mask) reinterpret_cast<float4*>(tl.mask)[q] = make_float4(m2[0].x, m2[0].y, m2[1].x, m2[1].y);
              if (tl.net_out) reinterpret_cast<float4*>(tl.net_out)[q] = make_float4(m1[0].x, m1[0].y, m1[1].x, m1[1].y);
              if (tl.net_raw) reinterpret_cast<float4*>(tl.net_raw)[q] = make_float4(raw[0].x, raw[0].y, raw[1].x, raw[1].y);
            }
          }
          if (++acc == kStripAcc) { acc = 0; accp ^= 1; }
        }
        continue;
      }
      // columns are processed in chunks of kChunk = min(kCols, 32); n_real divides kChunk, so column c always maps to
      // channel (c % kChunk) % n_real and the pooling partial sums need only kChunk registers
      constexpr int kChunk = kCols < 32 ? kCols : 32;
      constexpr int kSub = kEpi / 4;                      // warps per quadrant; warp `sub` owns chunks sub, sub + kSub, ...
      static_assert(kTail || (kCols / kChunk) % kSub == 0, "chunks must divide evenly among the warps of a quadrant");
      const int sub = (warp - kStripEpi0) >> 2;
      float pool_acc[kChunk];
#pragma unroll
      for (int c = 0; c < kChunk; ++c) pool_acc[c] = a.pool_max ? -INFINITY : 0.f;
      for (int j = un.j0; j < un.j1; ++j) {
        mbar_wait(smem_u32(&bars->acc_full[acc]), accp);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)kCols;
        const int64_t row0 = ((int64_t)un.b * a.out_h + (int64_t)j * a.up_h + G.ph0) * a.out_w + (int64_t)x * a.up_w;
#pragma unroll
        for (int cq = 0; cq < kCols / kChunk / kSub; ++cq) {
          const int ch = sub + cq * kSub;
          uint32_t rg[kChunk / 16][16];
#pragma unroll
          for (int cb = 0; cb < kChunk / 16; ++cb) tc_ld16(taddr + (uint32_t)(ch * kChunk + 16 * cb), rg[cb]);
          tc_ld_wait();
          if (cq == kCols / kChunk / kSub - 1) {   // this warp's share of the accumulator has been read: release it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
          }
#pragma unroll
          for (int s8 = 0; s8 < kChunk / 8; ++s8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int cl = 8 * s8 + q;
              v[q] = __uint_as_float(rg[cl / 16][cl % 16]) + bias_col[ch * kChunk + cl];
            }
            // activation selected once per group (a per-value runtime switch is ~8 instructions per accumulator element,
            // and the epilogue is bound by its own instruction stream)
            if (a.act == DCS_ACT_RELU) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
            } else if (a.act == DCS_ACT_LRELU) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.01f * v[q]);
            } else if (a.act == DCS_ACT_SIGMOID) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = sigmoidf_(v[q]);
            }
            if (valid) {
              if (a.pool_max) {
#pragma unroll
                for (int q = 0; q < 8; ++q) pool_acc[8 * s8 + q] = fmaxf(pool_acc[8 * s8 + q], v[q]);
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) pool_acc[8 * s8 + q] += v[q];
              }
            }
            if (valid) {
              const int c0 = ch * kChunk + 8 * s8, run = c0 >> a.run_log2, off = c0 & (run_len - 1);
              __nv_bfloat16* o = a.dst + (row0 + (int64_t)run * a.out_w) * a.n_real + off;
              uint32_t pk[4];
              if (a.f16) {
#pragma unroll
                for (int q = 0; q < 4; ++q) pk[q] = pack_f16x2(v[2 * q], v[2 * q + 1]);
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) pk[q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
              }
              *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        if (++acc == kStripAcc) { acc = 0; accp ^= 1; }
      }
      if (a.pool && a.pool_max) {   // per-(image, channel) maxima (the real path's AdaptiveMaxPool2d(1))
#pragma unroll
        for (int c = 0; c < kChunk; ++c) {
          float mx = pool_acc[c];
#pragma unroll
          for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          if (lane == 0) pool_max(a.pool + (int64_t)un.b * a.n_real + ((sub * kChunk + c) & (a.n_real - 1)), mx);
        }
      } else if (a.pool) {  // numerator of the ComplexAdaptiveAvgPool2d(1) that follows (c_network.py:208, 219)
        if constexpr (kChunk == 32) {
          const float sum = transpose_reduce32(pool_acc, lane);   // pool_acc[c] sums columns c, c + 32 kSub, ... of this warp
          pool_add(a.pool + (int64_t)un.b * a.n_real + ((sub * kChunk + lane) & (a.n_real - 1)), sum);   // chunk sub of the row
        } else {
#pragma unroll
          for (int c = 0; c < kChunk; ++c) {
            float sum = pool_acc[c];
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) pool_add(a.pool + (int64_t)un.b * a.n_real + (c & (a.n_real - 1)), sum);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace dcs

using namespace dcs;

static int ilog2_exact(int v) { return (v > 0 && (v & (v - 1)) == 0) ? __builtin_ctz(v) : -1; }

extern "C" int dcs_cconv2d_strip_fwd(const dcs_cstrip_params* p, void* stream) {
  DCS_REQUIRE(p && p->src0 && p->weights && p->items, "dcs_cconv2d_strip_fwd: null pointer");
  const dcs_strip_tail* tail = p->tail;
  DCS_REQUIRE(tail ? (tail->noisy_spec && tail->clean_spec) : (p->dst && p->bias), "dcs_cconv2d_strip_fwd: null output / bias pointer");
  DCS_REQUIRE(!tail || (p->cout == 1 && p->up_h == 2 && p->up_w == 8 && p->cols == 32 && p->out_w % 8 == 0 && !p->pool_sums),
              "dcs_cconv2d_strip_fwd: the tail epilogue is decoder[6] only (cout 1, 2 phase rows x 8 pixels per strip row)");
  DCS_REQUIRE(!tail || (tail->combine >= DCS_COMBINE_DCS && tail->combine <= DCS_COMBINE_DRS), "dcs_cconv2d_strip_fwd: bad combine mode");
  DCS_REQUIRE(p->batch > 0 && p->in_h > 0 && p->in_w > 0 && p->cout > 0, "dcs_cconv2d_strip_fwd: bad shape");
  DCS_REQUIRE(is_h16(p->dtype), "dcs_cconv2d_strip_fwd: dtype must be DCS_F16 or DCS_BF16");
  DCS_REQUIRE(p->stride_w == 1 || p->stride_w == 2, "dcs_cconv2d_strip_fwd: stride_w must be 1 or 2");
  DCS_REQUIRE(p->in_w % p->stride_w == 0, "dcs_cconv2d_strip_fwd: in_w must be a multiple of stride_w");
  DCS_REQUIRE(p->n_groups >= 1 && p->n_groups <= DCS_STRIP_MAX_GROUPS, "dcs_cconv2d_strip_fwd: bad n_groups");
  DCS_REQUIRE((p->c1 == 0) == (p->src1 == nullptr), "dcs_cconv2d_strip_fwd: src1 / c1 mismatch");
  const int P0 = 4 * p->c0 * p->stride_w, P1 = p->c1 ? 4 * p->c1 * p->stride_w : P0;   // strip row bytes (bf16 complex)
  DCS_REQUIRE((P0 == 32 || P0 == 64 || P0 == 128) && (P1 == 32 || P1 == 64 || P1 == 128),
              "dcs_cconv2d_strip_fwd: strip rows must be 32, 64 or 128 bytes (got %d, %d)", P0, P1);
  DCS_REQUIRE(p->box_units >= kStripM && p->box_units <= 256, "dcs_cconv2d_strip_fwd: box_units must be in [128, 256]");
  const int N = 2 * p->cout;
  DCS_REQUIRE(ilog2_exact(N) >= (tail ? 1 : 3), "dcs_cconv2d_strip_fwd: 2*cout must be a power of two >= 8");
  DCS_REQUIRE(p->cols == 32 || p->cols == 64 || p->cols == 128, "dcs_cconv2d_strip_fwd: cols must be 32, 64 or 128");
  DCS_REQUIRE(N <= 64, "dcs_cconv2d_strip_fwd: 2*cout must be <= 64 (wider layers use dcs_cconv2d_tc_fwd)");
  DCS_REQUIRE(N <= 32 || p->cols == N, "dcs_cconv2d_strip_fwd: 2*cout = 64 needs a single-phase accumulator (cols = 64)");
  DCS_REQUIRE(p->n_mma % 16 == 0 && p->n_mma >= 16 && p->n_mma <= p->cols, "dcs_cconv2d_strip_fwd: bad n_mma");
  const int run = p->up_w * N;
  DCS_REQUIRE(ilog2_exact(run) >= 3 && p->cols % run == 0, "dcs_cconv2d_strip_fwd: cols must be a multiple of up_w*2*cout");
  DCS_REQUIRE(((uintptr_t)p->src0 % 16 == 0) && ((uintptr_t)p->src1 % 16 == 0) && ((uintptr_t)p->weights % 16 == 0) &&
              ((uintptr_t)p->dst % 16 == 0), "dcs_cconv2d_strip_fwd: pointers must be 16-byte aligned");
  if (tail)
    DCS_REQUIRE(((uintptr_t)tail->noisy_spec % 16 == 0) && ((uintptr_t)tail->clean_spec % 16 == 0) && ((uintptr_t)tail->noise_spec % 16 == 0) &&
                ((uintptr_t)tail->mask % 16 == 0) && ((uintptr_t)tail->net_out % 16 == 0) && ((uintptr_t)tail->net_raw % 16 == 0),
                "dcs_cconv2d_strip_fwd: tail arrays must be 16-byte aligned");

  {  // the item table is read on the host (it is copied into the kernel parameters): refuse a device pointer
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p->items) == cudaSuccess)
      DCS_REQUIRE(at.type != cudaMemoryTypeDevice, "dcs_cconv2d_strip_fwd: items must be a HOST pointer");
    else
      cudaGetLastError();
  }
  StripArgs a;
  memset(&a, 0, sizeof(a));
  a.batch = p->batch;
  a.PH = p->out_h / p->up_h; a.PW = p->out_w / p->up_w;
  a.out_h = p->out_h; a.out_w = p->out_w; a.up_h = p->up_h; a.up_w = p->up_w; a.n_real = N; a.run_log2 = ilog2_exact(run);
  a.s_h = p->stride_h;
  a.n_strips = (a.PW + kStripM - 1) / kStripM;
  const uint32_t strip0 = ((uint32_t)(p->box_units * P0) + 1023u) & ~1023u;
  const uint32_t strip1 = p->c1 ? (((uint32_t)(p->box_units * P1) + 1023u) & ~1023u) : 0u;
  a.slot_bytes = strip0 + strip1; a.src1_off = strip0; a.has_src1 = p->c1 ? 1 : 0;
  a.tx_bytes = (uint32_t)(p->box_units * P0 + (p->c1 ? p->box_units * P1 : 0));
  a.row_bytes0 = (uint32_t)P0; a.row_bytes1 = (uint32_t)P1;
  a.n_mma = p->n_mma; a.act = p->act;
  a.f16 = p->dtype == DCS_F16 ? 1 : 0;
  a.pool_max = p->pool_mode == DCS_POOL_MAX ? 1 : 0;
  a.bias = p->bias; a.dst = reinterpret_cast<__nv_bfloat16*>(p->dst); a.pool = reinterpret_cast<long long*>(p->pool_sums);
  if (tail) a.tail = *tail;

  CUtensorMap tmA0, tmA1;
  const int w_units = p->in_w / p->stride_w;
  if (int e = make_act_map_generic(&tmA0, p->src0, a.f16, P0 / 2, w_units, p->in_h, p->batch, p->box_units)) return e;
  if (p->c1) { if (int e = make_act_map_generic(&tmA1, p->src1, a.f16, P1 / 2, w_units, p->in_h, p->batch, p->box_units)) return e; }
  else tmA1 = tmA0;
  cudaStream_t st = (cudaStream_t)stream;

  for (int g = 0; g < p->n_groups; ++g) {   // one launch per phase group (its weights stay resident in shared memory)
    const auto& s = p->group[g];
    DCS_REQUIRE(s.n_items > 0 && s.n_items <= kStripMaxItems && s.item0 >= 0 && s.item0 + s.n_items <= p->n_items_total,
                "dcs_cconv2d_strip_fwd: bad item range");
    DCS_REQUIRE(s.n_dy <= kStripMaxRing && s.n_dy >= p->stride_h && s.w_bytes > 0 && s.w_bytes % 16 == 0 && s.w_off % 16 == 0,
                "dcs_cconv2d_strip_fwd: bad group");
    DCS_REQUIRE(s.n_ph * run == p->cols, "dcs_cconv2d_strip_fwd: group phase rows x up_w x 2*cout must equal cols");
    DCS_REQUIRE(s.n_items % s.n_dy == 0, "dcs_cconv2d_strip_fwd: every ring row must carry the same number of items");
    const int ipr = s.n_items / s.n_dy;
    for (int i = 0; i < s.n_items; ++i)   // items sorted by drow, ipr per row: the kernel indexes them [drow][ipr]
      DCS_REQUIRE((int)p->items[s.item0 + i].drow == i / ipr, "dcs_cconv2d_strip_fwd: items must be sorted by drow, %d per ring row", ipr);
    a.grp.dy_min = s.dy_min; a.grp.n_dy = s.n_dy; a.grp.ph0 = s.ph0; a.grp.x_min = s.x_min; a.grp.w_bytes = (uint32_t)s.w_bytes;
    a.grp.w_ptr = reinterpret_cast<const unsigned char*>(p->weights) + s.w_off;
    memcpy(a.items, p->items + s.item0, (size_t)s.n_items * sizeof(uint4));
    for (int i = 0; i < s.n_items; ++i) {   // .w = high word of the item's A descriptor (SBO | version | swizzle of its source's rows)
      const uint32_t rb = (a.items[i].z >> 24) & 2u ? a.row_bytes1 : a.row_bytes0;
      const uint32_t layout = rb == 128 ? 2u : (rb == 64 ? 4u : 6u);
      a.items[i].w = ((8u * rb) >> 4) | (1u << 14) | (layout << 29);
    }
    a.w_smem_bytes = ((uint32_t)s.w_bytes + 1023u) & ~1023u;
    const size_t fixed = 1024 + a.w_smem_bytes + (size_t)p->cols * sizeof(float) + sizeof(StripBarriers);
    const size_t budget = 227 * 1024;
    DCS_REQUIRE(fixed + (size_t)(s.n_dy + p->stride_h) * a.slot_bytes <= budget,
                "dcs_cconv2d_strip_fwd: weights (%u B) + ring (%d x %u B) exceed shared memory", a.w_smem_bytes, s.n_dy + p->stride_h, a.slot_bytes);
    a.R = (int)std::min<size_t>(kStripMaxRing, (budget - fixed) / a.slot_bytes);
    const size_t smem = fixed + (size_t)a.R * a.slot_bytes;

    // row chunking: units = images x column strips x row chunks, statically striped over the CTAs.  A unit re-loads
    // (n_dy - s_h) halo rows, so pick the chunk count that minimises waves x (rows + halo cost).
    const int ctas = num_sms();
    double best = 1e30;
    for (int nc = 1; nc <= a.PH; ++nc) {
      const int rows = (a.PH + nc - 1) / nc;
      if ((a.PH + rows - 1) / rows != nc) continue;
      const int64_t units = (int64_t)p->batch * a.n_strips * nc;
      const int64_t waves = (units + ctas - 1) / ctas;
      const double cost = (double)waves * (rows + 0.5 * (s.n_dy - p->stride_h) + 1.0);
      if (cost < best) { best = cost; a.n_chunks = nc; a.chunk_rows = rows; }
    }
    a.n_units = p->batch * a.n_strips * a.n_chunks;
    const int grid = std::min(ctas, a.n_units);

    int threads = strip_threads(4);
#define DCS_STRIP_LAUNCH(...)                                                                                                     \
    do {                                                                                                                          \
      DCS_CUDA(cudaFuncSetAttribute(cconv_strip_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      cconv_strip_kernel<__VA_ARGS__><<<grid, threads, smem, st>>>(tmA0, tmA1, a);                                                \
    } while (0)
    // instantiated shapes (accumulator columns, ring rows per output row, MMA items per ring row)
    if (tail) {
      DCS_REQUIRE(s.n_dy == 3 && ipr == 12, "dcs_cconv2d_strip_fwd: no tail kernel instance for n_dy=%d items/row=%d", s.n_dy, ipr);
      threads = strip_threads(16);
      DCS_STRIP_LAUNCH(32, 3, 12, 16, true);                                            // decoder[6] + mask tail, 4-pixel strip rows
    }
    else if (p->cols == 32 && s.n_dy == 7 && ipr == 7) DCS_STRIP_LAUNCH(32, 7, 7);     // encoder[1]: k7 s(2,2), 1 source
    else if (p->cols == 64 && s.n_dy == 3 && ipr == 12) { threads = strip_threads(8); DCS_STRIP_LAUNCH(64, 3, 12, 8); }   // decoder[5] merged phases
    else if (p->cols == 64 && s.n_dy == 2 && ipr == 32) { threads = strip_threads(8); DCS_STRIP_LAUNCH(64, 2, 32, 8); }   // decoder[4], one phase row per launch
    else if (p->cols == 64 && s.n_dy == 2 && ipr == 24) { threads = strip_threads(8); DCS_STRIP_LAUNCH(64, 2, 24, 8); }   // decoder[4] merged pw
    else if (p->cols == 128 && s.n_dy == 7 && ipr == 4) { threads = strip_threads(16); DCS_STRIP_LAUNCH(128, 7, 4, 16); } // encoder[0]: Toeplitz blocks
    else if (p->cols == 64 && s.n_dy == 5 && ipr == 10) { threads = strip_threads(8); DCS_STRIP_LAUNCH(64, 5, 10, 8); }   // encoder[2]: k5 s(2,2), N = 64
    else DCS_REQUIRE(false, "dcs_cconv2d_strip_fwd: no kernel instance for cols=%d n_dy=%d items/row=%d", p->cols, s.n_dy, ipr);
#undef DCS_STRIP_LAUNCH
    DCS_LAUNCHED();
  }
  return 0;
}
