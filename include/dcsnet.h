/*
 * dcsnet.h — C ABI of libdcsnet_sm100a.so, the B200-native (sm_100a) replacement for the DCS-Net forward
 * hot path:  STFT -> complex encoder/decoder (C_NETWORK.forward) -> bounded mask / subtraction -> iSTFT.
 *
 * The reference (jackhwalters/DCS-Net) is pure Python/PyTorch and has NO FFI of its own; every GPU op on this
 * path is a torch-1.9 ATen dispatch (cuDNN / cuFFT / eltwise).  Each entry point below therefore cites the
 * reference *call site* whose ATen dispatches it replaces (file:line under /root/reference), and
 * INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer owned by the caller (PyTorch allocates);
 *     the library allocates nothing persistent on the device and never synchronises.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); entries are CUDA-graph capturable.
 *   - return 0 on success; non-zero = error (negative: argument/shape error, positive: cudaError_t);
 *     message via dcs_last_error_string() (thread-local).  There is no CPU fallback.
 *   - activation tensors are "channels-last complex": (B, H, W, C, 2) with the (re, im) pair innermost, which
 *     is the memory of a torch complex64 NCHW tensor in torch.channels_last format (dtype DCS_F32) or its
 *     16-bit twin (dtype DCS_F16 or DCS_BF16, used by the tcgen05 tensor-core modes).
 *   - spectrograms at the boundary use the reference layout (B, F=256, T) complex64, T contiguous.
 */
#ifndef DCSNET_H_
#define DCSNET_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCS_ABI_VERSION 2

/* activation storage types.  The tensor-core modes store activations in 16 bits: DCS_F16 (IEEE half, 11-bit significand;
 * the default — stores saturate at +-65504) or DCS_BF16 (8-bit significand, fp32 range).  kind::f16 MMAs take either. */
enum { DCS_F32 = 0, DCS_BF16 = 1, DCS_F16 = 2 };
/* Pooled sums (numerators of ComplexAdaptiveAvgPool2d(1): `pool_sums`, `sums`) are 64-bit fixed point, value * 2^28,
 * accumulated with integer atomics so that results are bit-identical from run to run. */
#define DCS_POOL_FRAC_BITS 28
/* pool_mode of the conv epilogues: sums (above) or per-(b, channel) MAXIMA (the real path's AdaptiveMaxPool2d(1)), stored as
 * the order-preserving unsigned image of the float + 1 (0 = empty, so the same memset-zero clear applies). */
enum { DCS_POOL_SUM = 0, DCS_POOL_MAX = 1 };
enum { DCS_ACT_NONE = 0, DCS_ACT_RELU = 1, DCS_ACT_LRELU = 2, DCS_ACT_SIGMOID = 3 }; /* ComplexReLU / ComplexLReLU(0.01) / ComplexSigmoid */
enum { DCS_COMBINE_DCS = 0, DCS_COMBINE_DC = 1,                  /* S = Y - Y*M   |   S = Y*M                (complex masks) */
       DCS_COMBINE_DR = 2, DCS_COMBINE_DRS = 3 };                 /* |S| = |Y| m   |   |S| = |Y| - |Y| m, noisy phase (real masks) */

#define DCS_MAX_TAPS 64

int dcs_abi_version(void);
const char* dcs_last_error_string(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t dcs_launch_count(void);

/* ---- a1: STFT front-end.  Replaces torch.stft(n_fft=512, hop=32, win=512, hann, normalized, center) [1:257]
 *      at data.py:112-134 (config.py:72-77).  audio (B, L) fp32 -> spec (B, 256, T) complex64, T = L/32 + 1.
 *      If bn_affine != NULL (6 floats A00 A01 A10 A11 c0 c1: the folded eval-mode `initial_batchnorm`,
 *      c_network.py:190) a second tensor bn_out (B,256,T,1) of dtype bn_dtype receives A*[re;im]+c; with bn_real != 0 the
 *      affine is applied to (|spec|, 0) instead — the real path's initial BatchNorm2d of the magnitude (r_network.py:128,
 *      network_functions.py:286), stored as the channel pair (bn(|Y|), padding). */
typedef struct {
  const float* audio; float* spec; int batch; int length; int n_frames;
  const float* bn_affine; void* bn_out; int bn_dtype; int bn_real;
} dcs_stft_params;
int dcs_stft_fwd(const dcs_stft_params* p, void* stream);

/* ---- f3: GPU data front-end = VoiceBankDataset.__getitem__ (data.py:68-143) for a batch:
 *      torchaudio Resample(48000 -> 16000) (config.py:61; 41-tap hann-windowed sinc, stride 3: `kernel` holds the taps,
 *      built on the host by the restated torchaudio formula), zero padding of short utterances, crop of `window`
 *      samples at start16[b] (16 kHz samples; NULL = 0), noise = noisy - clean, and the inf / nan checks:
 *      flags[b] |= 1 / 2 / 4 when clean / noisy / noise holds a non-finite value (flags may be NULL; zero it first).
 *      clean48 / noisy48: (B, stride48) fp32, valid lengths lengths48[b] (NULL = stride48).  Outputs (B, window) fp32;
 *      the three STFTs of data.py:115-134 are dcs_stft_fwd on them. */
typedef struct {
  const float* clean48; const float* noisy48; const int64_t* lengths48; const int64_t* start16;
  int batch; int64_t stride48; int window;
  const float* kernel; int n_taps; int orig; int width;
  float* clean16; float* noisy16; float* noise16; unsigned int* flags;
} dcs_frontend_params;
int dcs_frontend_fwd(const dcs_frontend_params* p, void* stream);

/* ---- a15: iSTFT back-end.  Replaces the polar round trip (network_functions.py:398-401) + mag_phase_2_wave
 *      (network_functions.py:140-150): abs / atan2(im, re+eps) / mag*cos / mag*sin, zero row appended at the END
 *      of the frequency axis, torch.istft(n_fft=512, hop=32, hann, normalized).  spec (B,256,T) complex64 ->
 *      audio (B, 32*(T-1)) fp32.  exact_polar=1 evaluates atan2f/cosf/sinf literally, 0 uses the algebraically
 *      identical (re+eps, im)/hypot form, 2 the same form with approximate reciprocal square roots (rel. error ~2e-7;
 *      the tensor-core mode's choice), 3 skips the round trip (spec is already mag * e^{j phase}: the real path's fused
 *      tail, which applies the NOISY phase itself, network_functions.py:300-304). */
typedef struct {
  const float* spec; float* audio; int batch; int n_frames; float atan2_eps; int exact_polar;
  /* mag_phase_2_wave(mag, phase, config) called directly (network_functions.py:140): if spec == NULL the input is
   * the pair of fp32 (B,256,T) arrays mag / phase and the kernel forms mag*cos(phase), mag*sin(phase) itself. */
  const float* mag; const float* phase;
} dcs_istft_params;
int dcs_istft_fwd(const dcs_istft_params* p, void* stream);

/* ---- a3: ComplexBatchNorm2d, eval mode, folded to a per-channel 2x2 affine (complexPyTorch 0.3; used at
 *      c_network.py:101,113,148) + optional activation (a5).  x,y channels-last complex, n_pix = B*H*W.
 *      affine: C x 6 floats (A00 A01 A10 A11 c0 c1). */
typedef struct {
  const void* x; void* y; const float* affine; int64_t n_pix; int channels; int act; int in_dtype; int out_dtype;
} dcs_cbn_params;
int dcs_cbn_apply(const dcs_cbn_params* p, void* stream);

/* ---- a4 / a12 / a8 / a11: complex convolution as ONE real implicit GEMM  (M = pixels, N = 2*Cout,
 *      K = taps * 2*Cin).  Replaces apply_complex over nn.Conv2d (encoder, c_network.py:107-112), over
 *      nn.ConvTranspose2d k3 s1 p1 (decoder, c_network.py:135-147; expressed as the flipped convolution), over
 *      nn.Linear (fc, c_network.py:202, as a 1x1 convolution), together with the skip torch.cat and
 *      complex_upsample that precede each decoder layer (c_network.py:214-216): the K loop walks the two
 *      sources src0 | src1 so the concatenation never exists, and nearest up-sampling is folded into
 *      `phases` sub-pixel classes with pre-summed taps.  Bias rule (b_r-b_i, b_r+b_i), the eval-mode BN
 *      affine and the activation are folded into `weight`/`bias`/`act` by the host-side packer.
 *
 *      Geometry: output pixel (oy, ox) = (j*up_h + ph, i*up_w + pw) for phase p = ph*up_w + pw and
 *      (j, i) in [0,out_h/up_h) x [0,out_w/up_w); tap t of phase p reads source pixel
 *      (j*stride_h + dy[p*ntaps+t], i*stride_w + dx[p*ntaps+t]), zero outside [0,in_h) x [0,in_w).
 *      weight: FFMA path  fp32 [phases][ntaps][2*(c0+c1)][2*cout]   (N contiguous)
 *              tcgen05    fp16 / bf16 (in_dtype F16 / BF16, kind::f16; same type as the activations) or fp32 pre-rounded
 *                         to tf32 (in_dtype F32, kind::tf32)
 *                         [phases][n_pad][K padded to 128 bytes]     (K contiguous), n_pad = max(16, 2*cout)
 *      bias:   fp32 [2*cout] added before `act`.
 *      pool_sums (optional, int64 fixed point [B][2*cout], pre-zeroed): per-(b, channel) sums of the epilogue output, i.e. the
 *      numerator of ComplexAdaptiveAvgPool2d(1) for the channel attention that follows (c_network.py:219). */
typedef struct {
  const void* src0; const void* src1; int c0; int c1;
  int batch; int in_h; int in_w;
  int out_h; int out_w; int cout;
  int up_h; int up_w; int stride_h; int stride_w;
  int ntaps; int8_t dy[DCS_MAX_TAPS]; int8_t dx[DCS_MAX_TAPS];
  const void* weight; const float* bias; int act;
  void* dst; int in_dtype; int out_dtype;
  int64_t* pool_sums;
  int pool_mode;   /* DCS_POOL_SUM / DCS_POOL_MAX (tcgen05 path) */
  int bias_phase_stride; /* tcgen05 path: floats between the bias vectors of consecutive phases; 0 = one bias vector for all phases */
} dcs_cconv_params;
int dcs_cconv2d_fwd(const dcs_cconv_params* p, void* stream);    /* fp32 CUDA-core path (<=1e-5 mode) */
int dcs_cconv2d_tc_fwd(const dcs_cconv_params* p, void* stream); /* tcgen05/TMEM/TMA path (fp16 / bf16 / tf32) */

/* ---- a4 / a12 / a11 (few-channel layers: encoder[0..2], decoder[4..6]): "row-strip" tensor-core convolution.
 *      Same math and reference call sites as dcs_cconv2d_tc_fwd (c_network.py:107-112, 135-147, 214-216); different
 *      data movement: one TMA box per SOURCE ROW (a strip of `box_units` strip rows; a strip row = one pixel, or a
 *      pixel pair when stride_w = 2, of 32 / 64 / 128 bytes) is loaded once into a ring of rows in shared memory and
 *      every tap that touches it is issued as a row-shifted / K-sliced UMMA descriptor over that strip; the layer's
 *      weights stay resident in shared memory.  The layer is described by a host-built table of MMA items
 *      (dcs-net_b200/packing.py: StripConv):
 *        a_off16  (strip_offset + shift*row_bytes + kslice*32) / 16, strip_offset = 0 for src0 and
 *                 round_up(box_units*row_bytes0, 1024) for src1
 *        b_off16  offset/16 of the item's [n_mma][16] bf16 weight block (rows of 32 B, SWIZZLE_32B image) in the group's
 *                 weight image
 *        d_col    first accumulator column; columns are ordered (phase row, phase col, n) so that one phase row is a
 *                 contiguous run of up_w*2*cout output elements
 *        drow     ring row relative to the first source row of the output row: source row = j*stride_h + dy_min + drow
 *        flags    bit 0: first MMA into its accumulator columns (overwrite), bit 1: reads src1
 *      Phase groups: group g owns output phase rows ph0 .. ph0+n_ph-1 of every up_h block and is served by its own
 *      CTAs (weights of one group resident per CTA).  16-bit storage only: `dtype` (DCS_F16 / DCS_BF16) is the type of
 *      src0 / src1 / weights / dst. */
#define DCS_STRIP_MAX_GROUPS 2
typedef struct { uint32_t a_off16; uint32_t b_off16; uint16_t d_col; uint8_t drow; uint8_t flags; uint32_t reserved; /* caller: 0; the library writes the item's A-descriptor high word into its own copy */ } dcs_strip_item;
typedef struct { int item0; int n_items; int dy_min; int n_dy; int ph0; int n_ph; int x_min; int w_bytes; int64_t w_off; } dcs_strip_group;
/* optional fused tail (decoder[6] only, replaces dcs_dec6_tail_fwd on the tensor-core path): the accumulator columns are
 * (phase row, 8 output pixels, re/im) of decoder[6]'s raw output; the epilogue adds (bias_re, bias_im) and applies
 * bound_cRM twice, the product with Y and the subtraction exactly as dcs_mask_combine does.  dst / bias are unused.
 * Real path (combine = DCS_COMBINE_DR / DRS; r_network.py:172 + network_functions.py:286-305, 338-342): the logit is the
 * .re column, m = sigmoid(logit) is written to `mask` as an fp32 REAL (B, 2h, 2w) array, and clean_spec / noise_spec
 * receive |Y| m (or |Y| - |Y| m) times e^{j atan2(Im Y, Re Y + eps)} as complex64, ready for dcs_istft_fwd with
 * exact_polar = 3; net_raw / net_out are ignored. */
typedef struct {
  const float* noisy_spec; float* net_raw; float* net_out; float* mask; float* noise_spec; float* clean_spec;
  float bias_re; float bias_im; float atan2_eps; int combine; int exact_polar;
} dcs_strip_tail;
typedef struct {
  const void* src0; const void* src1; int c0; int c1;
  int batch; int in_h; int in_w;
  int out_h; int out_w; int cout;
  int up_h; int up_w; int stride_h; int stride_w;
  int n_groups; dcs_strip_group group[DCS_STRIP_MAX_GROUPS];
  const dcs_strip_item* items; int n_items_total;   /* HOST pointer: the table is copied into the kernel parameters */
  const void* weights;
  int box_units; int n_mma; int cols;
  const float* bias; int act;
  void* dst; int64_t* pool_sums;
  const dcs_strip_tail* tail;   /* NULL: bias + activation + 16-bit store epilogue */
  int dtype;                    /* DCS_F16 or DCS_BF16 */
  int pool_mode;                /* DCS_POOL_SUM / DCS_POOL_MAX */
} dcs_cstrip_params;
int dcs_cconv2d_strip_fwd(const dcs_cstrip_params* p, void* stream);

/* ---- a9: ComplexChannelAttention (c_network.py:53-69; pools network_functions.py:114-138 — the "max" pool is
 *      an average pool, so the gate is sigmoid_c(2*fc(avg))).
 *      dcs_chan_pool: sums[b][c] (complex) = sum over H*W of x (channels-last).  sums must be pre-zeroed.
 *      dcs_chan_gate: gate[b][c] = sigmoid_c( 2 * W2 * crelu( W1 * (sums/hw) ) ), W1: (Cr,C) W2: (C,Cr) complex,
 *      given as separate real/imag fp32 matrices (the conv_r / conv_i 1x1 weights, no bias). */
typedef struct { const void* x; int64_t* sums; int batch; int hw; int channels; int dtype; } dcs_chan_pool_params;
int dcs_chan_pool(const dcs_chan_pool_params* p, void* stream);
/* same arguments; sums[b][c][re/im slot] = maximum over H*W in the DCS_POOL_MAX encoding (pre-zeroed) */
int dcs_chan_max(const dcs_chan_pool_params* p, void* stream);
/* mean[i] = sums[i] * 2^-DCS_POOL_FRAC_BITS * inv_hw  (n scalars): the pooled means themselves, i.e. ComplexAdaptiveAvgPool2d(1)
 * / ComplexAdaptiveMaxPool2d(1) called on their own (network_functions.py:114-138) */
int dcs_pool_mean(const int64_t* sums, float inv_hw, float* mean, int64_t n, void* stream);
typedef struct {
  const int64_t* sums; float inv_hw; float* gate; int batch; int channels; int reduced;
  const float* w1_r; const float* w1_i; const float* w2_r; const float* w2_i;
} dcs_chan_gate_params;
int dcs_chan_gate(const dcs_chan_gate_params* p, void* stream);

/* ---- a10: ComplexSpatialAttention (c_network.py:71-84) applied to u = gate_c * x (complex product, c_network.py
 *      :209,219).  dcs_spat_stats: stats[b][h][w] = { mean_c(u) (complex), max_c Re(u), max_c Im(u) } (4 floats).
 *      dcs_spat_apply: g = sigmoid_c( conv7x7_complex(stats) ), y = g * u (c_network.py:210-211, 220).
 *      chan_gate may be NULL (u = x); y may be NULL when only gate_out is wanted.
 *      w7: fp32 [2 (r,i)][2 (in ch: mean,max)][7][7]  = conv1.conv_r.weight, conv1.conv_i.weight. */
typedef struct {
  const void* x; const float* chan_gate; float* stats; int batch; int h; int w; int channels; int dtype;
  /* optional fused dcs_chan_gate: when sums != NULL the kernel computes the channel gate itself from the pooled sums
   * (sum over H*W of x) and the fc weights, ignores chan_gate, and writes the gate to gate_out (B, C) complex. */
  const int64_t* sums; int reduced; const float* w1_r; const float* w1_i; const float* w2_r; const float* w2_i; float* gate_out;
} dcs_spat_stats_params;
int dcs_spat_stats(const dcs_spat_stats_params* p, void* stream);
typedef struct {
  const void* x; const float* chan_gate; const float* stats; const float* w7; void* y;
  int batch; int h; int w; int channels; int in_dtype; int out_dtype;
  float* gate_out; /* optional (B,H,W) complex64: the spatial gate itself (ComplexSpatialAttention.forward's return) */
} dcs_spat_apply_params;
int dcs_spat_apply(const dcs_spat_apply_params* p, void* stream);

/* ---- a9 + a10 fused: y = SA(u) * u with u = CA(x) * x (c_network.py:208-211, 219-220) in ONE pass over x: the
 *      channel-gate MLP of dcs_chan_gate on the pooled `sums` (B,C) complex (sum over H*W of x, e.g. the conv
 *      epilogue's pool_sums), then per pixel tile: x tile + halo -> shared memory, statistics of dcs_spat_stats, the 7x7
 *      gate conv of dcs_spat_apply, and the product.  x is read once, y written once. */
typedef struct {
  const void* x; void* y; const int64_t* sums;
  int batch; int h; int w; int channels; int reduced; int in_dtype; int out_dtype;
  const float* w1_r; const float* w1_i; const float* w2_r; const float* w2_i; const float* w7;
  /* dcs_attention_stream only: real != 0 = the REAL CBAM of r_network.py:8-40 on a pair tensor (channels = C/2 pairs of
   * real channels): `sums` holds per-channel MAXIMA (DCS_POOL_MAX encoding), w1_r = fc.0.weight (R, C), w2_r = fc.2.weight
   * (C, R) (w1_i / w2_i unused), w7 = conv1.weight (2, 49); gate_c = sigmoid(W2 relu(W1 max)), statistics (mean_c u, max_c u),
   * gate_s = sigmoid(conv7x7), y = gate_s * gate_c * x element-wise.  fp16 storage. */
  int real;
} dcs_attention_params;
int dcs_attention_fused(const dcs_attention_params* p, void* stream);
/* Same contract, 16-bit storage only (tensor-core modes; in_dtype == out_dtype): streaming row-ring form — x read once by bulk copies, the 7x7 gate
 * conv (c_network.py:74,79-83) as mma.sync TF32 row-partials with a register ring of pending output rows, y written
 * once.  Replaces dcs_spat_stats + dcs_spat_apply on the tensor-core path (csrc/attention_stream.cu). */
int dcs_attention_stream(const dcs_attention_params* p, void* stream);

/* ---- f1 (real path): RealChannelAttention + RealSpatialAttention (r_network.py:8-40; applied at
 *      152-156, 163-165) on a channels-last REAL tensor x (B, H, W, C), fp32 or bf16 (`dtype`), C a power of two <= 256:
 *      gate_c = sigmoid(W2 relu(W1 maxpool_hw(x))) with w1 (R, C), w2 (C, R) (only the max-pool branch counts, line 24);
 *      u = gate_c * x; gate_s = sigmoid(conv7x7([mean_c u, max_c u])) with w7 = conv1.weight (1, 2, 7, 7) flattened;
 *      y = gate_s * u (same dtype as x).  workspace: dcs_real_attention_workspace_bytes(). */
typedef struct {
  const void* x; void* y; int batch; int h; int w; int channels; int reduced; int dtype;
  const float* w1; const float* w2; const float* w7;
  void* workspace; int64_t workspace_bytes;
} dcs_real_attention_params;
int64_t dcs_real_attention_workspace_bytes(int batch, int h, int w, int channels);
int dcs_real_attention_fwd(const dcs_real_attention_params* p, void* stream);

/* ---- f1 (real path): nn.LSTM(D -> 128, 2 layers, bidirectional, batch_first) of r_network.py:70-74 / 137-139.
 *      x (B, S, D) fp32 (the channels-last latent viewed as a sequence) -> y (B, S, 256) fp32.  Weights transposed for the
 *      kernels: w_ih0_t [D][2*4H] and w_ih1_t [2H][2*4H] (columns dir*4H + gate row, gate order i f g o),
 *      w_hh_t [layer][dir][H][4H], bias [layer][2*4H] = b_ih + b_hh (packing.PackedRNet.lstm_t). */
typedef struct {
  const void* x; float* y; int batch; int seq; int in_dim; int hidden; int in_dtype;
  const float* w_ih0_t; const float* w_ih1_t; const float* w_hh_t; const float* bias;
  void* workspace; int64_t workspace_bytes;
} dcs_rlstm_params;
int64_t dcs_rlstm_workspace_bytes(int batch, int seq, int hidden);
int dcs_rlstm_fwd(const dcs_rlstm_params* p, void* stream);

/* Tensor-core form of the same LSTM (fp16 / bf16 modes): x (B, S, D) and y (B, S, 256) in the 16-bit storage type `dtype`;
 *      input projections as tcgen05 kind::f16 GEMMs, recurrence as fp16 mma.sync with W_hh resident in registers.
 *      w_ih0: [dir 2][half 2][256][D] and w_ih1: [dir][half][256][2H] in `dtype`, K contiguous (gate rows half*256.. of the
 *      direction); w_hh: fp32 [layer][dir][4H][H] (the reference layout); bias: fp32 [layer][dir][4H] = b_ih + b_hh
 *      (packing.PackedRNet.lstm_tc). */
typedef struct {
  const void* x; void* y; int batch; int seq; int in_dim; int hidden; int dtype;
  const void* w_ih0; const void* w_ih1; const float* w_hh; const float* bias;
  void* workspace; int64_t workspace_bytes;
} dcs_rlstm_tc_params;
int64_t dcs_rlstm_tc_workspace_bytes(int batch, int seq, int hidden);
int dcs_rlstm_tc_fwd(const dcs_rlstm_tc_params* p, void* stream);

/* ---- a7: ComplexLSTM (c_network.py:12-51): real_lstm / imag_lstm = nn.LSTM(128->64, 2 layers, bidirectional),
 *      out = (R(re) - I(im)) + j (R(im) + I(re)).  x (B,S,D) complex channels-last (the latent, sequence index
 *      = h*W'+w, c_network.py:200) -> y (B,S,2*hidden) complex fp32.
 *      Weights: per lstm l in {real,imag}, layer in {0,1}, dir in {fwd,rev}: w_ih (4H, Din), w_hh (4H, H),
 *      bias (4H) = b_ih + b_hh; packed contiguously as [layer][lstm][dir].  workspace: see dcs_clstm_workspace_bytes. */
typedef struct {
  const void* x; float* y; int batch; int seq; int in_dim; int hidden; int in_dtype;
  const float* w_ih0; const float* w_ih1; const float* w_hh; const float* bias;
  void* workspace; int64_t workspace_bytes;
  /* optional tensor-core input projections (tf32 operands, kind::tf32): K-major weight slices [4 = lstm*2+dir][4H][D]
   * for layer 0 and [4][4H][2H] for layer 1, pre-rounded to tf32.  NULL -> fp32 CUDA-core GEMMs. */
  const float* w_ih0_t; const float* w_ih1_t;
  int seqs_per_cta; /* 0 = auto (4 when 2*batch % 4 == 0), or 2 */
  /* optional (tensor-core mode, with w_ih*_t): W_hh pre-rounded to tf32 in mma.sync m16n8k8 A-fragment order
   * [layer][lstm][dir][warp 8][tile 2][kstep 8][lane 32][4] (packing.pack_lstm); the recurrence then runs on the tensor
   * cores (TF32 products, fp32 accumulation).  NULL -> fp32 CUDA-core recurrence. */
  const float* w_hh_frag;
} dcs_clstm_params;
int64_t dcs_clstm_workspace_bytes(int batch, int seq, int hidden);
int dcs_clstm_fwd(const dcs_clstm_params* p, void* stream);

/* ---- a13 / a14: tail.  net_raw = decoder[6] output (B,256,T) complex64 (c_network.py:216,224);
 *      mask1 = bound_cRM(net_raw) (c_network.py:225) ; mask2 = bound_cRM(mask1) (network_functions.py:394);
 *      prod = Y (.) mask2 (complex_mat_mult, network_functions.py:90-96,396); dcs: S = Y - prod (397),
 *      dc: S = prod (434).  Any of net_out / mask / noise_spec may be NULL. */
typedef struct {
  const float* net_raw; const float* noisy_spec; float* net_out; float* mask; float* noise_spec; float* clean_spec;
  int64_t n; float atan2_eps; int combine; int exact_polar;
} dcs_mask_combine_params;
int dcs_mask_combine(const dcs_mask_combine_params* p, void* stream);

/* ---- a3 (initial_batchnorm) + a4 (encoder[0]) + a5 fused: ComplexConv2d(1 -> 8, k7, stride (2,2), pad 3) on
 *      initial_batchnorm(x) (c_network.py:190, 107-114), eval BN + ComplexReLU folded.  spec: (B, h, w) complex64
 *      (the spectrogram itself); bn_affine: 6 floats (folded initial_batchnorm) applied on load, padding stays 0;
 *      weight: fp32 [49 taps][2 (re,im of the input)][16] (same layout as the FFMA operand); dst (B, h/2, w/2, 8). */
typedef struct {
  const float* spec; const float* bn_affine; const float* weight; const float* bias;
  void* dst; int out_dtype; int batch; int h; int w;
} dcs_enc0_params;
int dcs_enc0_fwd(const dcs_enc0_params* p, void* stream);

/* ---- a12 (last layer) + a13 + a14 fused: decoder[6] = ComplexConvTranspose2d(16 -> 1, k3 s1 p1) applied to
 *      cat(d, skip_sa) up-sampled (2,2) (c_network.py:214-216, 134-140), then the whole tail of dcs_mask_combine.
 *      d, skip: (B, h, w, 8) channels-last complex of in_dtype; outputs (B, 2h, 2w) complex64.  N = 2 cannot feed a
 *      tensor core, so this is a CUDA-core kernel; decoder[6]'s raw output never reaches HBM (net_raw optional).
 *      weight: fp32 [4 phases][4 taps][16 ci][4] = the real 2x2 block of each pre-summed tap stored by columns
 *      (M00 M10 M01 M11), out.re = M00 x.re + M01 x.im, out.im = M10 x.re + M11 x.im. */
typedef struct {
  const void* d; const void* skip; int in_dtype; int batch; int h; int w;
  const float* weight; float bias_re; float bias_im;
  const float* noisy_spec; float* net_raw; float* net_out; float* mask; float* noise_spec; float* clean_spec;
  float atan2_eps; int combine; int exact_polar;
} dcs_dec6_tail_params;
int dcs_dec6_tail_fwd(const dcs_dec6_tail_params* p, void* stream);

/* ---- stand-alone element-wise functions of network_functions.py (used by the step functions outside forward):
 *      bound_cRM (77-88), complex_mat_mult (90-96), cRM (62-75, eps inside both denominators). n complex elements. */
int dcs_bound_crm(const float* x, float* y, int64_t n, float atan2_eps, int exact_polar, void* stream);
int dcs_cmul(const float* a, const float* b, float* y, int64_t n, void* stream);
int dcs_crm(const float* s, const float* y_noisy, float* m, int64_t n, float eps, void* stream);
/* real path (dr / drs) step functions: magnitude + phase = atan2(im, re + eps) of a complex64 array (network_functions.py:
 * 286-288; phase may be NULL) and the magnitude-mask combine (296-305: subtract = 1, clean = mag - mag*mask, noise = mag*mask;
 * 338-342: subtract = 0, clean = mag*mask).  mask element i is read at mask[i * mask_stride]. */
int dcs_mag_phase(const float* spec, float* mag, float* phase, int64_t n, float atan2_eps, void* stream);
int dcs_real_mask_combine(const float* mag, const float* mask, int64_t mask_stride, float* clean_mag, float* noise_mag,
                          int64_t n, int subtract, void* stream);

/* ---- a11 stand-alone: complex_upsample(mode='nearest') (complexPyTorch; c_network.py:215) on channels-last complex.
 *      (The fused path never materialises this: see dcs_cconv_params.up_h/up_w.) */
int dcs_upsample_nearest(const void* x, void* y, int batch, int h, int w, int channels, int up_h, int up_w, int dtype,
                         void* stream);

/* ==== f2: first kernels of the TRAINING step (network_functions.py:210-280 train_batch_2_loss,
 *      168-208 calc_loss; c_network.py:243-261 training_step).  Contracts in closed form: oracle/train_oracle.py. ==== */

/* ---- train-mode ComplexBatchNorm2d forward (complexPyTorch 0.3): batch statistics over n_pix pixels
 *      per channel (two-stage deterministic double-precision reduction), 2x2 inverse-square-root whitening, affine, optional
 *      activation; y = A x + c with the per-channel affine written to `affine` (C x 6, the dcs_cbn_apply operand).
 *      running_mean (C x 2: the complex64 buffer viewed as floats) / running_covar (C x 3) are updated in place with
 *      `momentum` (unbiased covariance, eps-inclusive diagonal, as the reference accumulates them); num_batches_tracked
 *      (int64, optional) += 1; saved (C x 8, optional): mean.re, mean.im, Rrr, Rii, Rri, Crr, Cii, Cri for the backward. */
typedef struct {
  const void* x; void* y; int64_t n_pix; int channels; int act; int in_dtype; int out_dtype;
  const float* weight; const float* bias; float eps; float momentum;
  float* running_mean; float* running_covar; int64_t* num_batches_tracked;
  float* affine; float* saved; void* workspace; int64_t workspace_bytes;
} dcs_cbn_train_params;
int64_t dcs_cbn_train_workspace_bytes(int64_t n_pix, int channels);
int dcs_cbn_train_fwd(const dcs_cbn_train_params* p, void* stream);
/* ---- its backward (no activation): x, dy fp32 channels-last complex, `saved` from the forward ->
 *      dx, dweight (C x 3), dbias (C x 2).  Pass 1 reduces eight per-channel sums, a per-channel 3x3 Jacobian of the whitening
 *      matrix gives the coefficients of pass 2, dx = P dy + Q x + k. */
typedef struct {
  const void* x; const float* dy; float* dx; int64_t n_pix; int channels;
  const float* saved; const float* weight; float* dweight; float* dbias; void* workspace; int64_t workspace_bytes;
  float* conv_bias_grad_r; float* conv_bias_grad_i;   /* optional (both or neither), C floats each: conv_r / conv_i .bias.grad of the convolution
                                                         in front of this BatchNorm = per-channel sums of dx (S.re + S.im, S.im - S.re) */
  int x_dtype;                                        /* storage type of x (the saved forward input): DCS_F32 / DCS_F16 / DCS_BF16; dy, dx fp32 */
} dcs_cbn_train_bwd_params;
int dcs_cbn_train_bwd(const dcs_cbn_train_bwd_params* p, void* stream);

/* ---- SiSNR (network_functions.py:30-42) per batch row: value[row] = 10 log10(|s_t|^2 / (|e - s_t|^2 + eps) + eps) and
 *      grad = grad_scale / rows * d value / d estimate (the loss is the batch mean; calc_loss's signs and alpha go into
 *      grad_scale).  value or grad may be NULL. */
int dcs_si_snr(const float* clean, const float* estimate, int rows, int length, float eps, float grad_scale, float* value,
               float* grad, void* stream);
/* ---- adjoint of mag_phase_2_wave's iSTFT (network_functions.py:140-150): waveform gradient (B, 32 (T-1)) -> spectrogram
 *      gradient (B, 256, T) complex64 in the dL/dRe + j dL/dIm convention (the STFT kernel in adjoint mode). */
int dcs_istft_adjoint(const float* grad_audio, float* grad_spec, int batch, int n_frames, void* stream);
/* ---- adjoint of the fused mask tail (dcs_mask_combine / the dec6 tail): iSTFT-adjoint gradients of the clean (and, dcs,
 *      noise: g_noise != NULL) waveforms -> gradient w.r.t. decoder[6]'s raw output: polar^T, combine^T (dM = conj(Y) dN),
 *      bound_cRM^T twice.  n complex elements. */
int dcs_mask_tail_bwd(const float* net_raw, const float* noisy_spec, const float* g_clean, const float* g_noise, float* d_raw,
                      int64_t n, float atan2_eps, void* stream);
/* ---- adjoint of torch.cat((d, skip), 1) + complex_upsample (c_network.py:214-215): g (B, h*up_h, w*up_w, c0 + c1) = the
 *      conv dgrad of the up-sampled concatenation -> gd (B, h, w, c0), gskip (B, h, w, c1), fp32 channels-last complex.
 *      (The dgrad itself is dcs_cconv2d_tc_fwd / dcs_cconv2d_fwd with the role-swapped weights of packing.dgrad_conv.) */
int dcs_upcat_adjoint(const float* g, float* gd, float* gskip, int batch, int h, int w, int c0, int c1, int up_h, int up_w,
                      void* stream);

/* ---- weight gradient of ComplexConv2d on the tensor cores (csrc/wgrad_tc.cu; oracle/train_oracle.cconv2d_backward):
 *      x (B, in_h, in_w, cin) and dy (B, out_h, out_w, cout) channels-last complex in 16-bit storage `dtype`;
 *      dWp[n][tap][k] = sum_pixels dy[pix][n] x[pix * stride + off(tap)][k] as ONE tcgen05 GEMM with K = pixels (both operands
 *      MN-major straight from the activations' layout, TMA-staged, zero padding by out-of-bounds fill), split over K and folded:
 *      dw_r, dw_i fp32 (cout, cin, ntaps) = the reference's conv_r.weight.grad / conv_i.weight.grad with tap = ky*kw + kx,
 *      offsets dy_off[tap] = ky - pad_h, dx_off[tap] = kx - pad_w.  Needs 2*cout % 128 == 0, 2*cin in {64, 128, 192, 256}. */
typedef struct {
  const void* x; const void* dy; int dtype;
  int batch; int in_h; int in_w; int out_h; int out_w; int cin; int cout; int stride_h; int stride_w;
  int ntaps; int8_t dy_off[DCS_MAX_TAPS]; int8_t dx_off[DCS_MAX_TAPS];
  float* dw_r; float* dw_i; void* workspace; int64_t workspace_bytes;
} dcs_cwgrad_params;
int64_t dcs_cwgrad_workspace_bytes(const dcs_cwgrad_params* p);
int dcs_cwgrad_tc(const dcs_cwgrad_params* p, void* stream);

/* ==== f2, second slice (csrc/train_bwd.cu): the rest of the backward pass, fp32, deterministic two-stage reductions.
 *      Contracts: oracle/train_oracle.py (cconv2d_backward, decoder_stage_backward, clinear_backward, attention_backward,
 *      lstm_forward_saved / lstm_bptt); reference: c_network.py:12-51, 53-84, 187-240; network_functions.py:210-280. ==== */

/* ---- generic weight-gradient implicit GEMM: dwp[tap][k][n] = sum over output pixels (b, oh, ow) of
 *      x[(b, oh*stride_h + dy_off[tap], ow*stride_w + dx_off[tap]) * x_pitch + k] * dy[(b, oh, ow) * dy_pitch + n]
 *      (zero outside the input), k < k2, n < n2 REAL channels.  Covers every ComplexConv2d / ComplexConvTranspose2d /
 *      ComplexLinear weight gradient (x = the layer's real-block input, k2 = 2 cin, n2 = 2 cout; fold with
 *      dcs_wgrad_fold_complex) and the LSTM's dW_ih / dW_hh (one tap; h shifted by one step = dx_off -1 / +1 with in_w = S). */
typedef struct {
  const float* x; const float* dy;
  int batch; int in_h; int in_w; int out_h; int out_w; int k2; int n2; int x_pitch; int dy_pitch; int stride_h; int stride_w;
  int ntaps; int8_t dy_off[DCS_MAX_TAPS]; int8_t dx_off[DCS_MAX_TAPS];
  float* dwp; void* workspace; int64_t workspace_bytes;
} dcs_wgrad_params;
int64_t dcs_wgrad_workspace_bytes(const dcs_wgrad_params* p);
int dcs_wgrad(const dcs_wgrad_params* p, void* stream);
/* dwp [ntaps][2 cin][2 cout] -> conv_r.weight.grad / conv_i.weight.grad: dw_r = dWp[re,re] + dWp[im,im], dw_i = dWp[re->im] -
 * dWp[im->re]; layout (cout, cin, ntaps), or for ComplexConvTranspose2d (transposed = 1; cin / cout = those of the EQUIVALENT
 * conv the forward runs) the module's (cout_conv = its in_channels ... ) tensor (cin, cout, ntaps) with the taps reversed. */
int dcs_wgrad_fold_complex(const float* dwp, int ntaps, int cin, int cout, int transposed, float* dw_r, float* dw_i, void* stream);
/* dst[c * rows + r] = src[r * src_pitch + c]: dwp[k][n] -> a real weight's .grad (n, k) */
int dcs_transpose(const float* src, float* dst, int rows, int cols, int src_pitch, void* stream);
/* C[m][n] = sum_k A[m * lda + k] * B[n * ldb_n + k * ldb_k] (+ bias[n]) (+ C[m][n] when accumulate): the LSTM projections
 * (NT: ldb_k = 1) and their data gradients (NN: ldb_n = 1) in the training step */
int dcs_sgemm(const float* A, int lda, const float* B, int ldb_n, int ldb_k, const float* bias, float* C, int ldc, int M, int N, int K,
              int accumulate, void* stream);
/* column sums of x (rows x cols, row pitch `pitch`): mode 0: out0[c] = S[c] (out1, optional, receives a copy: bias_ih / bias_hh);
 * mode 1 (complex conv / linear bias, columns (co, re / im)): out0[co] = S[2co] + S[2co+1] = conv_r.bias.grad,
 * out1[co] = S[2co+1] - S[2co] = conv_i.bias.grad */
int64_t dcs_colsum_workspace_bytes(int64_t rows, int cols);
int dcs_colsum(const float* x, int64_t rows, int cols, int pitch, int mode, float* out0, float* out1, void* workspace,
               int64_t workspace_bytes, void* stream);
/* zero insertion for the data gradient of a STRIDED conv: out (B, in_h, in_w, channels) complex = dy (B, out_h, out_w, channels)
 * placed at (oh * stride_h, ow * stride_w), zero elsewhere; the dgrad is then the stride-1 forward conv kernel with the flipped,
 * in / out-swapped weights (train_ops.dgrad_conv) */
int dcs_dilate(const float* dy, float* out, int batch, int out_h, int out_w, int in_h, int in_w, int channels, int stride_h,
               int stride_w, void* stream);
/* data gradient of a ComplexConv2d with ONE input channel (encoder[0], c_network.py:107-112) from the raw conv_r / conv_i weights
 * (cout, 1, kh, kw): dx (B, in_h, in_w) complex = sum over the taps of this pixel's stride phase (padding k // 2) */
int dcs_cconv_dgrad_cin1(const float* dy, const float* w_r, const float* w_i, float* dx, int batch, int in_h, int in_w, int out_h, int out_w,
                         int cout, int kh, int kw, int stride_h, int stride_w, void* stream);
/* z (B, h*up_h, w*up_w, c0 + c1) = complex_upsample(cat(d, skip)) (c_network.py:214-215), materialised as the wgrad's x operand in
 * out_dtype (fp32 for dcs_wgrad, fp16 / bf16 for dcs_wgrad_tc16) */
int dcs_upcat_fwd(const void* d, const void* skip, int in_dtype, void* z, int out_dtype, int batch, int h, int w, int c0, int c1, int up_h, int up_w,
                  void* stream);
/* decoder[6] backward in one kernel (c_network.py:214-216: ComplexConvTranspose2d(c0 + c1 -> 1, k3 s1 p1) on the (2,2) nearest
 * up-sampling of cat(d, skip)): dpre (B, 2h, 2w) complex = gradient of the layer's output -> g_d (B, h, w, c0), g_skip (B, h, w, c1),
 * the weight gradients in the module's (c0 + c1, 1, 3, 3) layout and the two bias gradients.  c0 + c1 <= 16. */
int64_t dcs_dec6_bwd_workspace_bytes(void);
int dcs_dec6_bwd(const void* d, const void* skip, int in_dtype, const float* dpre, const float* w_r, const float* w_i, int batch, int h, int w, int c0, int c1,
                 float* g_d, float* g_skip, float* dw_r, float* dw_i, float* db_r, float* db_i, void* workspace, int64_t workspace_bytes,
                 void* stream);
/* dz = act'(y) (.) (g0 + g1 + chan_const[b][c]) per real component (ComplexReLU / ComplexLReLU act on the parts); y = the
 * activation's OUTPUT (B, hw, channels) complex in storage type y_dtype (the saved forward tensor; gradients are fp32); g1 (same shape) and
 * chan_const (B, channels) complex are optional */
int dcs_act_bwd(const void* y, int y_dtype, const float* g0, const float* g1, const float* chan_const, float* dz, int batch, int64_t hw, int channels,
                int act, void* stream);
/* torch.nn.Dropout(p) on n floats (c_network.py:195-196 / 203-204 / 221-222: applied on view_as_real): y = x * keep / (1 - p),
 * keep drawn from Philox4x32-10 with key `seed` and counter offset + i / 4: the same (seed, offset) regenerates the mask for the
 * backward.  (The reference draws from torch's generator; the streams differ, the distribution does not.) */
int dcs_dropout(const void* x, void* y, int64_t n, int dtype, float p, uint64_t seed, uint64_t offset, void* stream);

/* ---- backward of the attended product y = s (.) (a (.) x) (ComplexChannelAttention + ComplexSpatialAttention applied with
 *      complex products, c_network.py:208-211 / 219-220).  Inputs saved by the forward: x, the channel gate a (B, C), the
 *      per-pixel statistics (B, hw, 4), the spatial gate s (B, hw) (dcs_spat_apply's gate_out), the pooled sums.
 *      dx = conj(a) du WITHOUT the per-channel constant davg / hw, which is returned in chan_const (B, C) and added by the
 *      consumer (dcs_act_bwd).  Weight gradients in the reference's layouts: fc.0 (R, C), fc.2 (C, R), conv1 (1, 2, 7, 7). */
typedef struct {
  const void* x; const float* dy; const float* gate_c; const float* stats; const float* gate_s; const float* w7; const void* sums;
  int batch; int h; int w; int channels; int reduced;
  const float* w1_r; const float* w1_i; const float* w2_r; const float* w2_i;
  float* dspre; float* dx; float* chan_const;
  float* dw1_r; float* dw1_i; float* dw2_r; float* dw2_i; float* dw7_r; float* dw7_i;
  void* workspace; int64_t workspace_bytes;
  int x_dtype;     /* storage type of x (DCS_F32 / DCS_F16 / DCS_BF16); dy, dx and the saved gates / statistics are fp32 */
} dcs_attention_bwd_params;
int64_t dcs_attention_bwd_workspace_bytes(int batch, int h, int w, int channels, int reduced);
int dcs_attention_bwd(const dcs_attention_bwd_params* p, void* stream);

/* ---- LSTM in training form (one direction pair of one layer; hidden = 64): pre (Q, S, 2 dirs, 4H) = x W_ih^T + b_ih + b_hh,
 *      w_hh (n_groups, 2, 4H, H) with sequence q using group q / (Q / n_groups); outputs h (Q, S, 2, H) and, for BPTT, the
 *      activated gates (Q, S, 2, 4H) and cell states (Q, S, 2, H).  dcs_lstm_train_bwd runs the reverse-time recurrence:
 *      dh (Q, S, 2, H) -> dpre (Q, S, 2, 4H); dW_ih, dW_hh, db, dx are GEMMs / reductions outside it. */
int dcs_lstm_train_fwd(const float* pre, const float* w_hh, int n_seq, int n_groups, int steps, int hidden, float* h, float* gates,
                       float* cells, void* stream);
int dcs_lstm_train_bwd(const float* w_hh, const float* gates, const float* cells, const float* dh, int n_seq, int n_groups, int steps,
                       int hidden, float* dpre, void* stream);
/* (n) interleaved complex <-> two planes (2, n); ComplexLSTM's combine (c_network.py:39-47) on h (2 lstm, 2 part, n):
 * out = (R(re) - I(im)) + j (R(im) + I(re)), and its adjoint */
int dcs_cplx_split(const void* x, int in_dtype, float* planes, int64_t n, void* stream);
int dcs_cplx_merge(const float* planes, float* x, int64_t n, void* stream);
int dcs_clstm_combine(const float* h, float* out, int64_t n, void* stream);
int dcs_clstm_combine_bwd(const float* dout, float* dh, int64_t n, void* stream);

/* ---- optimizer (c_network.py:229-235: torch.optim.Adam(lr, eps, weight_decay, amsgrad=True); config.py:48-49 clip by global
 *      norm 100): sum of squares of a flat fp32 buffer into a device double (accumulate = add to *out), then one fused update:
 *      g = grad * grad_scale * min(1, max_norm / (grad_scale * sqrt(*grad_sumsq) + 1e-6)) + weight_decay * p, m / v / vmax / p. */
int dcs_sumsq(const float* x, int64_t n, double* out, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);
int dcs_adam_amsgrad(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, const double* grad_sumsq, float max_norm,
                     float grad_scale, void* stream);
/* dst[i] = sum_j sign4[4i + j] * src[idx4[4i + j]] (idx < 0 = no term), stored as out_dtype: raw parameters -> kernel operand
 * layouts (block matrices, phase pre-sums, role swaps) by an index table built once on the host, one launch per step;
 * out_dtype: DCS_F32 / DCS_F16 / DCS_BF16, or 3 = fp32 rounded to tf32 (the kind::tf32 operands) */
int dcs_gather_pack(const float* src, const int32_t* idx4, const int8_t* sign4, void* dst, int64_t n, int out_dtype, void* stream);

/* ---- dcs_wgrad on the tensor cores: the tcgen05 GEMM of dcs_cwgrad_tc (K = pixels, MN-major operands straight from 16-bit
 *      channels-last activations) behind dcs_wgrad's interface — real channel counts k2 <= 256 (wider inputs: call per channel slice,
 *      x + offset with x_pitch = channels per pixel, dwp + k_offset * n2 with dwp_tap_stride = k2_total * n2), any n2; operand boxes
 *      beyond the real channels are zero-filled by TMA.  dwp [ntaps][k2][n2] fp32 as dcs_wgrad writes it. */
typedef struct {
  const void* x; const void* dy; int dtype;
  int batch; int in_h; int in_w; int out_h; int out_w; int k2; int n2; int x_pitch; int dy_pitch; int stride_h; int stride_w;
  int ntaps; int8_t dy_off[DCS_MAX_TAPS]; int8_t dx_off[DCS_MAX_TAPS];
  float* dwp; int64_t dwp_tap_stride; void* workspace; int64_t workspace_bytes;
} dcs_wgrad16_params;
int64_t dcs_wgrad_tc16_workspace_bytes(const dcs_wgrad16_params* p);
int dcs_wgrad_tc16(const dcs_wgrad16_params* p, void* stream);

/* ---- developer aid: per-CTA wait-cycle counters of the tcgen05 kernel (8 uint64 per CTA, >= 148 CTAs); NULL = off */
int dcs_tc_set_debug_buffer(void* dev_ptr);

/* ---- layout helpers for the layer-wise drop-in modules: fp32 <-> fp16 / bf16 copies of channels-last activations */
int dcs_convert(const void* src, void* dst, int64_t n_floats, int in_dtype, int out_dtype, void* stream);
/* zero `bytes` bytes of device memory on `stream` (cudaMemsetAsync: a memset node when captured, not a kernel): the
 * per-step clear of the pooled-sum accumulators */
int dcs_zero(void* dst, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCSNET_H_ */
