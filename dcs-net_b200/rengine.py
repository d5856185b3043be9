"""RealForwardPlan — the real-valued DR-Net / DRS-Net inference path (BASELINE configs[2]) on the tensor
cores: STFT -> |Y|, BatchNorm2d -> R_NETWORK.forward -> sigmoid mask -> |S| = |Y| m (dr) or |Y| - |Y| m (drs) -> iSTFT with the
NOISY phase, as a fixed sequence of sm_100a kernels over pre-allocated HBM buffers, replayable as a CUDA graph.  It follows
/root/reference/r_network.py:125-173 and network_functions.py:286-305 (drs) / 338-342 (dr) step by step (eval mode).

The real network runs on the complex path's kernels through PAIR PACKING (packing.PackedRNet): real channels (2c, 2c+1) are the
(re, im) of pseudo-complex channel c, so a real weight matrix is the block matrix the implicit-GEMM kernels multiply by and the
channel counts (16, 32, 64, 128, 256, 256, 256 real = 8 ... 128 pairs) give exactly the complex network's tile geometry:
  * every Conv2d / ConvTranspose2d (+ eval BatchNorm2d + ReLU / LeakyReLU, cat + nearest up-sampling folded in) on
    cconv_strip_kernel / cconv_tc_kernel (tcgen05, 16-bit storage, fp32 accumulation);
  * nn.LSTM(256 -> 128, 2 layers, bidirectional): dcs_rlstm_tc_fwd (kind::f16 projections, fp16 mma.sync recurrence);
  * the real CBAM (max-pool channel gate, mean / max spatial statistics, 7x7 gate conv): dcs_real_attention_fwd;
  * decoder[6] + sigmoid + magnitude combine + noisy phase fused into the strip kernel's tail epilogue; the iSTFT kernel then
    reads |S| e^{j phase} directly (exact_polar = 3).
Modes: 'fp16' (default) / 'bf16' as in engine.py.  (The fp32 <= 1e-5 mode of the real path is R_NETWORK.forward's CUDA-core
kernel sequence in r_network.py.)
"""
import os

import torch

from . import _lib as L
from . import ops, packing
from .engine import build_strips, MODES, TC_MODE, UPSAMPLE


class PackedRealNet(packing.PackedRNet):
    """PackedRNet + the row-strip operands, for one device and tensor-core mode."""

    def __init__(self, model_or_sd, device, mode=TC_MODE, no_of_layers=7):
        mode = TC_MODE if mode == "tc" else mode
        assert mode in MODES and mode != "fp32", "RealForwardPlan is the tensor-core path (fp16 / bf16); fp32 = R_NETWORK.forward"
        self.mode, self.device, self.tc = mode, device, True
        self.act_dtype = torch.float16 if mode == "fp16" else torch.bfloat16
        super().__init__(model_or_sd, device=device, no_of_layers=no_of_layers, tc_dtype=self.act_dtype)
        self.strip = build_strips(self.enc, self.dec, no_of_layers, device)


class RealForwardPlan:
    def __init__(self, packed, batch, n_frames, n_bins=256, variant="drs", atan2_eps=10e-7, exact_polar=False, keep_taps=False,
                 want_aux=True):
        if not torch.cuda.is_available():
            raise RuntimeError("dcsnet_b200.RealForwardPlan needs a CUDA device (sm_100a); there is no CPU fallback")
        L.lib()
        assert variant in ("dr", "drs")
        if n_frames % 8 or n_bins % 128:
            raise ValueError(f"R_NETWORK needs T % 8 == 0 and F % 128 == 0 (got F={n_bins}, T={n_frames}); "
                             "the reference fails at the skip torch.cat (r_network.py:160)")
        self.pk, self.B, self.T, self.F = packed, batch, n_frames, n_bins
        self.variant, self.eps, self.exact = variant, float(atan2_eps), bool(exact_polar)
        self.keep_taps, self.want_aux = keep_taps, want_aux
        dev, Lr = packed.device, packed.L
        self.device = torch.device(dev)
        adt = self.adt = packed.act_dtype
        B, T, F = batch, n_frames, n_bins
        new = lambda *s, dtype=adt: torch.empty(*s, dtype=dtype, device=dev)   # noqa: E731
        self.Y = new(B, F, T, dtype=torch.complex64)
        self.bn0 = new(B, F, T, 1, 2)
        self.enc = []
        H, W = F, T
        for i in range(Lr):
            H, W = ops.conv_out_hw(packed.enc[i], H, W)
            self.enc.append(new(B, H, W, packed.enc[i].cout, 2))
        self.S = H * W
        hid2 = packed.lstm_tc["w_hh"].shape[3] * 2                        # 2 * hidden = LSTM output features
        self.lat = new(B, H, W, hid2 // 2, 2)
        self.lstm_ws = torch.empty(ops.rlstm_tc_workspace_bytes(B, self.S), dtype=torch.uint8, device=dev)
        self.fc = new(B, H, W, packed.fc.cout, 2)
        self.skip, self.dec, self.datt = [], [], []
        att_bytes = 0
        for i in range(Lr):
            src = self.enc[Lr - 1 - i]
            self.skip.append(new(*src.shape))
            H, W = src.shape[1] * UPSAMPLE[i][0], src.shape[2] * UPSAMPLE[i][1]
            last = i == Lr - 1
            self.dec.append(None if last else new(B, H, W, packed.dec[i].cout, 2))
            self.datt.append(None if last else new(B, H, W, packed.dec[i].cout, 2))
            for t in (src, None if last else self.dec[i]):
                if t is not None:
                    att_bytes = max(att_bytes, int(L.lib().dcs_real_attention_workspace_bytes(B, t.shape[1], t.shape[2], 2 * t.shape[3])))
        self.att_ws = torch.empty(att_bytes, dtype=torch.uint8, device=dev)
        # fp16: streaming single-pass attention (attention_stream.cu, REAL variant) fed by per-(image, channel) maxima that
        # the producing conv's epilogue accumulates (DCS_POOL_MAX); one int64 buffer for all tensors, cleared once per step
        self.stream_attention = adt == torch.float16 and os.environ.get("DCS_STREAM_ATTENTION", "1") != "0"
        chans = [t.shape[3] for t in self.enc] + [t.shape[3] for t in self.dec[:-1]]
        self.pool_all = new(B * 2 * sum(chans), dtype=torch.int64)
        views, off = [], 0
        for c in chans:
            views.append(self.pool_all[off:off + B * c * 2].view(B, c, 2))
            off += B * c * 2
        self.pool_enc, self.pool_dec = views[:len(self.enc)], views[len(self.enc):]
        self.clean_spec = new(B, F, T, dtype=torch.complex64)
        self.noise_spec = new(B, F, T, dtype=torch.complex64) if (want_aux and variant == "drs") else None
        self.mask = new(B, F, T, dtype=torch.float32) if want_aux else None
        self.audio_in = new(B, 32 * (T - 1), dtype=torch.float32)
        self.audio_out = new(B, 32 * (T - 1), dtype=torch.float32)
        self.noise_audio = new(B, 32 * (T - 1), dtype=torch.float32) if self.noise_spec is not None else None
        self.graph = None
        self.taps = {}

    # ------------------------------------------------------------------ building blocks
    def _attention(self, x, att, y, maxima=None):
        """y = gate_s * gate_c * x (r_network.py:152-156 / 163-165) on the pair tensor viewed as 2C real channels.
        `maxima`: the per-channel maxima accumulated by the kernel that produced x (None: the three-pass kernels)."""
        w12, w7 = att
        if maxima is not None:
            return ops.real_attention_stream(x, maxima, w12, w7, y)
        return ops.real_attention(x, w12, w7, y=y, workspace=self.att_ws)

    def _conv(self, pk, src0, src1, dst, strip=None, pool=None):
        """Returns True when the kernel accumulated the channel maxima of dst into `pool`."""
        pool = pool if self.stream_attention else None
        if strip is not None and src0.shape[2] % pk.stride[1] == 0:
            ops.cconv_strip(strip, src0, src1, dst, pool_sums=pool, pool_max=True)
            return pool is not None
        use_tc = (2 * pk.cin) % 16 == 0 and pk.w_tc is not None
        ops.cconv(pk, src0, src1, dst, use_tc=use_tc, pool_sums=pool if use_tc else None, pool_max=True)
        if pool is not None and not use_tc:
            ops.chan_max(dst, pool)
        return pool is not None

    def _tap(self, name, t):
        if self.keep_taps:
            self.taps[name] = t

    # ------------------------------------------------------------------ the kernel sequence
    def _network(self):
        """bn0 -> ... -> (d, skip) in front of decoder[6] (r_network.py:128-165)."""
        pk, Lr = self.pk, self.pk.L
        strip0 = pk.strip.get(("enc", 0)) if self.T % 16 == 0 else None
        if self.stream_attention:
            ops.zero_(self.pool_all)
        x = self.bn0
        pooled_e = [False] * Lr
        for i in range(Lr):
            if i == 0 and strip0 is not None:
                pool = self.pool_enc[0] if self.stream_attention else None
                ops.cconv_strip(strip0, packing.StripEnc0.view_src(self.bn0), None, self.enc[0], pool_sums=pool, pool_max=True)
                pooled_e[0] = pool is not None
            else:
                pooled_e[i] = self._conv(pk.enc[i], x, None, self.enc[i], pk.strip.get(("enc", i)) if i else None, self.pool_enc[i])
            x = self.enc[i]
            self._tap(f"enc{i}", x)
        B, H, W, Cp, _ = x.shape
        ops.rlstm_tc(x.view(B, H * W, 2 * Cp), pk.lstm_tc, self.lat.view(B, H * W, -1), self.lstm_ws)   # sequence index = h * W + w
        self._tap("lstm", self.lat)
        self._conv(pk.fc, self.lat.view(B, 1, H * W, -1, 2), None, self.fc.view(B, 1, H * W, -1, 2))
        d = self.fc
        self._tap("fc", d)
        for i in range(Lr):
            e = Lr - 1 - i
            skip = self._attention(self.enc[e], pk.skip_att[i], self.skip[i], self.pool_enc[e] if pooled_e[e] else None)
            self._tap(f"skip{i}", skip)
            if i == Lr - 1:
                return d, skip      # decoder[6] is fused with the mask tail
            pooled = self._conv(pk.dec[i], d, skip, self.dec[i], pk.strip.get(("dec", i)), self.pool_dec[i])
            d = self._attention(self.dec[i], pk.dec_att[i], self.datt[i], self.pool_dec[i] if pooled else None)
            self._tap(f"dec{i}", d)

    def _tail(self, d_skip):
        d, skip = d_skip
        combine = L.COMBINE_DRS if self.variant == "drs" else L.COMBINE_DR
        strip6 = self.pk.strip.get(("dec", 6)) if d.shape[2] % 4 == 0 else None
        if strip6 is None:
            raise NotImplementedError("RealForwardPlan: decoder[6] needs the strip kernel (T % 8 == 0 guarantees it for the default geometry)")
        ops.dec6_tail_strip(strip6, d, skip, self.Y, self.clean_spec, mask=self.mask, noise_spec=self.noise_spec,
                            atan2_eps=self.eps, combine=combine, exact_polar=self.exact)

    def _enqueue_from_audio(self, with_noise_audio=False):
        # the STFT kernel also emits BatchNorm2d(|Y|) as the pair tensor encoder[0] reads by TMA
        ops.stft(self.audio_in, self.Y, bn_affine=self.pk.bn0, bn_out=self.bn0, bn_real=True)
        self._tail(self._network())
        ops.istft(self.clean_spec, self.audio_out, self.eps, 3)          # 3: the tail already applied the noisy phase
        if with_noise_audio and self.noise_audio is not None:
            ops.istft(self.noise_spec, self.noise_audio, self.eps, 3)

    def _enqueue_from_spec(self):
        """Y already holds the noisy spectrogram: |Y| -> BatchNorm2d -> network -> tail (no iSTFT)."""
        mag, _ = ops.mag_phase(self.Y, self.eps, want_phase=False)
        pair = torch.view_as_complex(torch.stack([mag, torch.zeros_like(mag)], dim=-1))     # layout glue: (|Y|, padding slot)
        ops.cbn_apply(torch.view_as_real(pair).view(self.B, self.F, self.T, 1, 2), self.pk.bn0, self.bn0)
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
        """audio (B, 32(T-1)) fp32 on device (or already in self.audio_in) -> enhanced audio (view of plan buffer)."""
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
        return dict(mask=self.mask, noise_spec=self.noise_spec, clean_spec=self.clean_spec)
