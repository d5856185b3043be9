"""ForwardPlan — the fused DCS-Net inference path: STFT -> C_NETWORK.forward -> bound_cRM x2 -> mask (.) Y
[-> Y - N] -> iSTFT, as a fixed sequence of sm_100a kernels over pre-allocated HBM buffers, replayable as a
CUDA graph.  It follows /root/reference/c_network.py:187-226 and network_functions.py:393-401 step by step
(eval mode: dropout is the identity, BN uses running statistics).

Memory plan (per plan instance, B utterances of T frames, everything resident in HBM):
  Y          (B,256,T)      complex64   noisy spectrogram (reference layout, T contiguous)
  bn0        (B,256,T,1,2)  act dtype   initial_batchnorm output = encoder input
  enc[i]     (B,H_i,W_i,C_i,2) act      encoder outputs, channels-last complex
  lat / fc   (B,2,T/8,128,2)            ComplexLSTM output (fp32) and ComplexLinear output (act dtype)
  skip[i], dec[i], datt[i]              attended skips, decoder conv outputs, attended decoder outputs
  raw        (B,256,T) complex64        decoder[6] output;   S / N / M / net_out outputs, audio (B, 32(T-1))
Modes (PackedNet(mode=...)):
  'fp32' : act dtype float32, CUDA-core FFMA GEMMs                                  -> S within 1e-5 of the reference
  'fp16' : act dtype float16, tcgen05 kind::f16 GEMMs with fp32 accumulation        -> S within 2e-3 (measured ~5e-4)
           THE tensor-core mode ('tc' is an alias).  fp16 carries tf32's 11-bit significand; stores saturate at 65504.
  'bf16' : act dtype bfloat16, same kernels and speed.  8-bit significands put S at ~3.6e-3 on the randomised-BN parity
           state (27 roundings between the STFT and the mask) — outside the 2e-3 budget; kept for checkpoints whose
           activations exceed fp16's range.
Attention statistics, LSTM state, masks, spectrograms and audio are fp32 in every mode.  Pooled sums are int64 fixed
point, so a step is bit-reproducible.
"""
import os

import torch

from . import _lib as L
from . import ops, packing

KERNEL_E = [7, 7, 5, 5, 3, 3, 3]                                            # config.py:83
STRIDE_E = [(2, 2), (2, 2), (2, 2), (2, 1), (2, 1), (2, 1), (2, 1)]          # config.py:99
UPSAMPLE = [(2, 1), (2, 1), (2, 1), (2, 1), (2, 2), (2, 2), (2, 2)]          # config.py:105


MODES = ("fp32", "fp16", "bf16")
TC_MODE = "fp16"          # the tensor-core mode the benchmark is quoted on ('tc')


def _check_model_geometry(model):
    """The kernel plan is built for the reference's default geometry (config.py:83-105, BN eps 1e-5).  A model built
    with another stride / up-sampling / kernel table or BN eps is refused here instead of being computed silently with the
    default geometry."""
    cfg = getattr(model, "config", None)
    if cfg is not None:
        for name, want in (("kernel_sizeE", KERNEL_E), ("strideE", STRIDE_E), ("upsample_scale_factor", UPSAMPLE)):
            got = getattr(cfg, name, None)
            if got is not None and [tuple(g) if isinstance(g, (list, tuple)) else g for g in got][:len(want)] != \
                    [tuple(w) if isinstance(w, (list, tuple)) else w for w in want][:len(got)]:
                raise NotImplementedError(f"dcsnet_b200 kernels are built for config.{name} = {want}, got {got}")
    if hasattr(model, "modules"):
        for m in model.modules():
            eps = getattr(m, "eps", None)
            if eps is not None and type(m).__name__.endswith("BatchNorm2d") and abs(eps - packing.BN_EPS) > 1e-12:
                raise NotImplementedError(f"dcsnet_b200 folds BatchNorm with eps = {packing.BN_EPS}, got {eps} in {type(m).__name__}")


def _sd_tensor_dict(model_or_sd):
    sd = model_or_sd.state_dict() if hasattr(model_or_sd, "state_dict") else model_or_sd
    return {k: v.detach() for k, v in sd.items()}


def build_strips(enc, dec, Lr, device):
    """Row-strip operands (packing.StripConv / StripEnc0 / StripDec6) for the few-channel layers of a packed encoder /
    decoder stack — complex net or the real net in pair packing (same geometry): {(kind, layer): strip}."""
    strip = {}
    if os.environ.get("DCS_STRIP", "1") == "0":
        return strip
    e0 = enc[0]
    if (e0.cin, e0.cout, e0.kh, e0.kw, tuple(e0.stride)) == (1, 8, 7, 7, (2, 2)) and os.environ.get("DCS_STRIP_ENC0", "1") != "0":
        strip[("enc", 0)] = packing.StripEnc0(e0, device=device)
    d6 = dec[Lr - 1] if Lr == 7 else None
    if d6 is not None and (d6.cin, d6.cout, tuple(d6.up)) == (16, 1, (2, 2)) and os.environ.get("DCS_STRIP_DEC6", "1") != "0":
        strip[("dec", 6)] = packing.StripDec6(d6, device=device)
    want = {("enc", 1): (8, 0, True, 1), ("dec", 4): (32, 32, False, 2), ("dec", 5): (16, 16, True, 1)}
    if os.environ.get("DCS_STRIP_ENC2", "1") != "0":
        want[("enc", 2)] = (16, 0, True, 1)       # k5 s(2,2), N = 64: 50 MMA items, 100 KB of resident weights
    for (kind, i), (c0, c1, merged, groups) in want.items():
        pc = (enc if kind == "enc" else dec)[i] if i < Lr else None
        if pc is not None and pc.cin == c0 + c1 and (2 * pc.cout) in ((16, 32, 64) if (kind, i) == ("enc", 2) else (16, 32)):
            strip[(kind, i)] = packing.StripConv(pc, c0, c1, merged=merged, groups=groups, device=device)
    return strip


class PackedNet:
    """All GEMM-ready operands of a C_NETWORK state_dict for one device and mode."""

    def __init__(self, model_or_sd, device, mode="fp32", no_of_layers=7):
        _check_model_geometry(model_or_sd)
        sd = _sd_tensor_dict(model_or_sd)
        mode = TC_MODE if mode == "tc" else mode
        assert mode in MODES, f"mode must be one of {MODES + ('tc',)}"
        self.mode, self.device, self.L = mode, device, no_of_layers
        self.tc = mode in ("fp16", "bf16")
        self.act_dtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[mode]
        tcd = self.act_dtype if self.tc else None
        bf = self.tc
        Lr = no_of_layers
        self.bn0 = packing.affine6(*packing.bn_affine_from_sd(sd, "initial_batchnorm.")).to(device)
        self.enc, self.dec, self.skip_ca, self.skip_sa, self.dec_ca, self.dec_sa = [], [], [], [], [], []
        for i in range(Lr):
            p = f"encoder.{i}.0."
            self.enc.append(packing.PackedConv(
                sd[p + "conv_r.weight"], sd[p + "conv_i.weight"], sd[p + "conv_r.bias"], sd[p + "conv_i.bias"],
                bn=packing.bn_affine_from_sd(sd, f"encoder.{i}.1."), stride=STRIDE_E[i], act=L.ACT_RELU,
                device=device, tc_dtype=tcd))
        for i in range(Lr):
            last = i == Lr - 1
            p = f"decoder.{i}." if last else f"decoder.{i}.0."
            self.dec.append(packing.PackedConv(
                sd[p + "conv_tran_r.weight"], sd[p + "conv_tran_i.weight"], sd[p + "conv_tran_r.bias"],
                sd[p + "conv_tran_i.bias"], bn=None if last else packing.bn_affine_from_sd(sd, f"decoder.{i}.1."),
                transposed=True, up=UPSAMPLE[i], act=L.ACT_NONE if last else L.ACT_LRELU, device=device, tc_dtype=tcd))
            self.skip_ca.append(packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", device))
            self.skip_sa.append(packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", device))
            if not last:  # decoder_attention[12], [13] never run (c_network.py:218)
                self.dec_ca.append(packing.pack_channel_attention(sd, f"decoder_attention.{2 * i}.", device))
                self.dec_sa.append(packing.pack_spatial_attention(sd, f"decoder_attention.{2 * i + 1}.", device))
        # few-channel layers: row-strip tensor-core kernel (csrc/cconv_strip.cu)
        self.strip = build_strips(self.enc, self.dec, Lr, device) if bf else {}
        self.lstm = packing.pack_lstm(sd, "lstm.", device)
        w_r, w_i = sd["fc.fc_r.weight"], sd["fc.fc_i.weight"]
        self.fc = packing.PackedConv(w_r[:, :, None, None], w_i[:, :, None, None], sd["fc.fc_r.bias"], sd["fc.fc_i.bias"],
                                     device=device, tc_dtype=tcd, want_tf32=bf)


class ForwardPlan:
    def __init__(self, packed, batch, n_frames, n_bins=256, variant="dcs", atan2_eps=10e-7, exact_polar=False,
                 keep_taps=False, want_aux=True):
        if not torch.cuda.is_available():
            raise RuntimeError("dcsnet_b200.ForwardPlan needs a CUDA device (sm_100a); there is no CPU fallback")
        L.lib()
        assert variant in ("dcs", "dc")
        if n_frames % 8 or n_bins % 128:
            raise ValueError(f"C_NETWORK needs T % 8 == 0 and F % 128 == 0 (got F={n_bins}, T={n_frames}); "
                             "the reference fails at the skip torch.cat (c_network.py:214)")
        self.pk, self.B, self.T, self.F = packed, batch, n_frames, n_bins
        self.variant, self.eps, self.exact = variant, float(atan2_eps), bool(exact_polar)
        self.keep_taps, self.want_aux = keep_taps, want_aux
        dev, Lr = packed.device, packed.L
        self.tc = packed.tc
        adt = packed.act_dtype
        self.adt = adt
        self.device = torch.device(dev)
        B, T, F = batch, n_frames, n_bins
        new = lambda *s, dtype=adt: torch.empty(*s, dtype=dtype, device=dev)
        self.Y = new(B, F, T, dtype=torch.complex64)
        self.bn0 = new(B, F, T, 1, 2)
        self.enc = []
        H, W = F, T
        for i in range(Lr):
            H, W = ops.conv_out_hw(packed.enc[i], H, W)
            self.enc.append(new(B, H, W, packed.enc[i].cout, 2))
        S = H * W
        self.S = S
        self.lat = new(B, H, W, 128, 2, dtype=torch.float32)
        self.lstm_ws = torch.empty(ops.clstm_workspace_bytes(B, S) // 4, dtype=torch.float32, device=dev)
        self.fc = new(B, H, W, packed.fc.cout, 2)
        self.skip, self.dec, self.datt = [], [], []
        max_c, max_hw = 128, 0
        for i in range(Lr):
            src = self.enc[Lr - 1 - i]
            self.skip.append(new(*src.shape))
            H, W = src.shape[1] * UPSAMPLE[i][0], src.shape[2] * UPSAMPLE[i][1]
            last = i == Lr - 1
            self.dec.append(new(B, H, W, packed.dec[i].cout, 2, dtype=torch.float32 if last else adt))
            self.datt.append(None if last else new(B, H, W, packed.dec[i].cout, 2))
            max_hw = max(max_hw, src.shape[1] * src.shape[2], 0 if last else H * W)
        # per-tensor pooling accumulators (numerators of ComplexAdaptiveAvgPool2d(1)), zeroed once per step; the
        # tcgen05 conv epilogue accumulates into them so the attended tensors are not re-read for the pooling
        self.fuse_pool = self.tc
        chans = [t.shape[3] for t in self.enc] + [t.shape[3] for t in self.dec[:-1]]
        self.pool_all = new(B * 2 * sum(chans), dtype=torch.int64)   # fixed point (include/dcsnet.h: DCS_POOL_FRAC_BITS)
        views, off = [], 0
        for c in chans:
            views.append(self.pool_all[off:off + B * c * 2].view(B, c, 2))
            off += B * c * 2
        self.pool_enc, self.pool_dec = views[:len(self.enc)], views[len(self.enc):]
        self.sums = new(B, max_c, 2, dtype=torch.int64)
        self.gate = new(B, max_c, 2, dtype=torch.float32)
        self.stats = new(B, max_hw, 4, dtype=torch.float32)
        self.clean_spec = new(B, F, T, dtype=torch.complex64)
        self.noise_spec = new(B, F, T, dtype=torch.complex64) if (want_aux and variant == "dcs") else None
        self.mask = new(B, F, T, dtype=torch.complex64) if want_aux else None
        self.net_out = new(B, F, T, dtype=torch.complex64) if want_aux else None
        self.audio_in = new(B, 32 * (T - 1), dtype=torch.float32)
        self.audio_out = new(B, 32 * (T - 1), dtype=torch.float32)
        self.noise_audio = new(B, 32 * (T - 1), dtype=torch.float32) if self.noise_spec is not None else None
        self.graph = None
        self.taps = {}
        # the skip attentions depend only on encoder outputs, so they can run on a side stream concurrently with the
        # ComplexLSTM + fc (fork / join is captured into the CUDA graph as parallel branches).  Measured on B200 at
        # batch 64 x 4 s: no gain (9.73 vs 9.72 ms) — the recurrence is issue-bound on 128 SMs, so it stays off.
        self.overlap = os.environ.get("DCS_OVERLAP", "0") == "1"
        # channel gate + spatial statistics + 7x7 gate conv + product in one kernel per attended tensor
        # (opt-in with DCS_FUSED_ATTENTION=1: measured slower than the three separate kernels on B200 — its load / stats / conv /
        # apply phases serialise inside a CTA; kept for the next round's TMA-pipelined version)
        self.fused_attention = os.environ.get("DCS_FUSED_ATTENTION", "0") == "1"
        # bf16 mode: streaming row-ring attention (x read once, TF32 tensor-core gate conv; csrc/attention_stream.cu)
        self.stream_attention = self.tc and os.environ.get("DCS_STREAM_ATTENTION", "1") != "0"
        self.side = torch.cuda.Stream(device=dev)

    # ------------------------------------------------------------------ building blocks
    def _attention(self, x, ca, sa_w7, y, sums=None):
        """y = SA(CA(x) * x) * (CA(x) * x)   (c_network.py:208-211 / 219-220).  `sums`: pooling numerators already
        accumulated by the kernel that produced x; otherwise they are computed here."""
        B, H, W, Cn, _ = x.shape
        gate = self.gate.view(-1)[: B * Cn * 2].view(B, Cn, 2)
        stats = self.stats.view(-1)[: B * H * W * 4].view(B, H * W, 4)
        if sums is None:
            sums = self.sums.view(-1)[: B * Cn * 2].view(B, Cn, 2)
            ops.zero_(sums)
            ops.chan_pool(x, sums)
        if self.stream_attention and Cn >= 8 and x.dtype in ops.H16 and y.dtype == x.dtype:
            return ops.attention_stream(x, sums, ca, sa_w7, y)
        if self.fused_attention and Cn >= 4:
            return ops.attention_fused(x, sums, ca, sa_w7, y)
        ops.spat_stats(x, None, stats, sums=sums, ca=ca, gate_out=gate)   # channel-gate MLP fused into the statistics pass
        ops.spat_apply(x, gate, stats, sa_w7, y)
        return y

    def _conv(self, pk, src0, src1, dst, pool=None, strip=None):
        """Returns (dst, pooled): pooled is True when the kernel accumulated the pooling sums into `pool`."""
        if strip is not None and self.tc and src0.dtype in ops.H16 and src0.shape[2] % pk.stride[1] == 0:
            fused = self.fuse_pool and pool is not None
            ops.cconv_strip(strip, src0, src1, dst, pool_sums=pool if fused else None)
            return dst, fused
        use_tc = self.tc and (2 * pk.cin) % 16 == 0 and (
            (src0.dtype in ops.H16 and pk.w_tc is not None) or (src0.dtype == torch.float32 and pk.w_tc32 is not None))
        fused = use_tc and self.fuse_pool and pool is not None
        ops.cconv(pk, src0, src1, dst, use_tc=use_tc, pool_sums=pool if fused else None)
        return dst, fused

    def _tap(self, name, t):
        if self.keep_taps:
            self.taps[name] = t

    # ------------------------------------------------------------------ the kernel sequence
    def _network(self, bn0_ready=False):
        """bn0 -> ... -> decoder[6] raw output (c_network.py:193-222)."""
        pk, Lr = self.pk, self.pk.L
        e0 = pk.enc[0]
        fused_first = (e0.cin, e0.cout, e0.kh, e0.kw, e0.stride) == (1, 8, 7, 7, (2, 2))
        # encoder[0] on the tensor cores (row-strip kernel, Toeplitz blocks) reads the bf16 initial_batchnorm output
        strip0 = pk.strip.get(("enc", 0)) if (self.tc and self.T % 16 == 0) else None
        if (self.keep_taps or not fused_first or strip0 is not None) and not bn0_ready:
            ops.cbn_apply(torch.view_as_real(self.Y).view(self.B, self.F, self.T, 1, 2), pk.bn0, self.bn0)
        x = self.bn0
        if self.fuse_pool:
            ops.zero_(self.pool_all)
        enc_pooled = [False] * Lr

        def skip_attention(i):
            e = Lr - 1 - i
            return self._attention(self.enc[e], pk.skip_ca[i], pk.skip_sa[i], self.skip[i],
                                   sums=self.pool_enc[e] if enc_pooled[e] else None)

        # A skip attention depends only on its encoder output: run it right behind the producing conv while that tensor is
        # still in the 126 MB L2 (tensors that fit), instead of re-reading it from HBM in the decoder loop.
        mode = os.environ.get("DCS_EARLY_SKIP", "auto")
        limit = {"0": -1, "1": 1 << 62}.get(mode, 80 << 20)
        early = lambda e: self.enc[e].numel() * self.enc[e].element_size() <= limit
        for i in range(Lr):
            if i == 0 and strip0 is not None:
                x = ops.cconv_strip(strip0, packing.StripEnc0.view_src(self.bn0), None, self.enc[0],
                                    pool_sums=self.pool_enc[0] if self.fuse_pool else None)
                enc_pooled[0] = self.fuse_pool
            elif i == 0 and fused_first:  # initial_batchnorm + encoder[0] straight from the spectrogram
                x = ops.enc0(e0, self.Y, pk.bn0, self.enc[0])
            else:
                x, enc_pooled[i] = self._conv(pk.enc[i], x, None, self.enc[i], self.pool_enc[i], pk.strip.get(("enc", i)))
            self._tap(f"enc{i}", x)
            if early(i):
                skip_attention(Lr - 1 - i)
        B, H, W, _, _ = x.shape

        main = torch.cuda.current_stream(self.device)
        # The attentions of the large (not L2-resident) skip tensors depend only on encoder outputs: with `overlap` they
        # run on a side stream next to the latency-bound ComplexLSTM recurrence (one CTA per SM, ~30 % of the issue
        # slots) and join before their decoder stage.  Needs the streaming attention (no shared scratch buffers).
        late = [i for i in range(Lr) if not early(Lr - 1 - i)]
        side_run = self.overlap and self.stream_attention and bool(late)
        if side_run:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                for i in late:
                    skip_attention(i)
        ops.clstm(x.view(B, H * W, x.shape[3], 2), self.lat.view(B, H * W, 128, 2), pk.lstm, self.lstm_ws, use_tc=self.tc,
                  seqs_per_cta=0)
        self._tap("lstm", self.lat)
        self._conv(pk.fc, self.lat.view(B, 1, H * W, 128, 2), None, self.fc.view(B, 1, H * W, 128, 2))
        d = self.fc
        self._tap("fc", d)
        joined = False
        for i in range(Lr):
            if early(Lr - 1 - i):
                skip = self.skip[i]
            elif side_run:
                if not joined:
                    main.wait_stream(self.side)
                    joined = True
                skip = self.skip[i]
            else:
                skip = skip_attention(i)
            self._tap(f"skip{i}", skip)
            if i == Lr - 1:
                return d, skip  # decoder[6] is fused with the mask tail (dcs_dec6_tail_fwd)
            d, pooled = self._conv(pk.dec[i], d, skip, self.dec[i], self.pool_dec[i], pk.strip.get(("dec", i)))
            self._tap(f"dec{i}_act", d)
            d = self._attention(d, pk.dec_ca[i], pk.dec_sa[i], self.datt[i], sums=self.pool_dec[i] if pooled else None)
            self._tap(f"dec{i}", d)

    def _tail(self, d_skip):
        d, skip = d_skip
        combine = L.COMBINE_DCS if self.variant == "dcs" else L.COMBINE_DC
        pk6 = self.pk.dec[self.pk.L - 1]
        strip6 = self.pk.strip.get(("dec", 6)) if (self.tc and d.dtype in ops.H16 and d.shape[2] % 4 == 0) else None
        if strip6 is not None:
            raw = self.dec[-1] if self.keep_taps else None
            ops.dec6_tail_strip(strip6, d, skip, self.Y, self.clean_spec, net_raw=raw, net_out=self.net_out, mask=self.mask,
                                noise_spec=self.noise_spec, atan2_eps=self.eps, combine=combine, exact_polar=self.exact)
            if self.keep_taps:
                self._tap(f"dec{self.pk.L - 1}", raw)
        elif pk6.w_tail is not None:
            raw = self.dec[-1] if self.keep_taps else None
            ops.dec6_tail(pk6, d, skip, self.Y, self.clean_spec, net_raw=raw, net_out=self.net_out, mask=self.mask,
                          noise_spec=self.noise_spec, atan2_eps=self.eps, combine=combine, exact_polar=self.exact)
            if self.keep_taps:
                self._tap(f"dec{self.pk.L - 1}", raw)
        else:  # non-default channel configuration: un-fused last layer + tail
            raw, _ = self._conv(pk6, d, skip, self.dec[-1])
            ops.mask_combine(raw, self.Y, self.clean_spec, net_out=self.net_out, mask=self.mask,
                             noise_spec=self.noise_spec, atan2_eps=self.eps, combine=combine, exact_polar=self.exact)

    def _enqueue_from_audio(self, with_noise_audio=False):
        strip0 = self.tc and self.T % 16 == 0 and ("enc", 0) in self.pk.strip
        if strip0:  # the STFT kernel also emits the folded initial_batchnorm output (bf16) that encoder[0] reads by TMA
            ops.stft(self.audio_in, self.Y, bn_affine=self.pk.bn0, bn_out=self.bn0)
        else:
            ops.stft(self.audio_in, self.Y)
        self._tail(self._network(bn0_ready=strip0))
        polar = 1 if self.exact else (2 if self.tc else 0)   # tensor-core mode: MUFU-only polar round trip (rel. ~2e-7)
        ops.istft(self.clean_spec, self.audio_out, self.eps, polar)
        if with_noise_audio and self.noise_audio is not None:
            ops.istft(self.noise_spec, self.noise_audio, self.eps, polar)

    def _enqueue_from_spec(self):
        self._tail(self._network())

    # ------------------------------------------------------------------ public
    def capture(self):
        """Capture the audio->audio pipeline into a CUDA graph (one launch per step afterwards)."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue_from_audio()  # warm-up: cudaFuncSetAttribute, lazy module load
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            before = L.launch_count()
            with torch.cuda.graph(g, stream=s):       # an explicit capture stream ON THE PLAN'S DEVICE (torch's default capture stream is
                # created once per process on whatever device was current then: a plan on another GPU captured an empty graph)
                self._enqueue_from_audio()
            self.graph_launches = L.launch_count() - before
            self.graph = g
        return self

    def enhance_audio(self, audio=None):
        """audio (B, 32(T-1)) fp32 on device (or already in self.audio_in) -> enhanced audio (view of plan buffer).
        Runs on the plan's device whatever the caller's current device is."""
        with torch.cuda.device(self.device):
            if audio is not None:
                self.audio_in.copy_(audio, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._enqueue_from_audio()
        return self.audio_out

    def enhance_spec(self, spec):
        """noisy spectrogram (B,F,T) complex64 -> dict of spectrogram-domain outputs (views of plan buffers)."""
        with torch.cuda.device(self.device):
            self.Y.copy_(spec)
            self._enqueue_from_spec()
        return dict(net_out=self.net_out, mask=self.mask, noise_spec=self.noise_spec, clean_spec=self.clean_spec)
