"""TrainStep — the GPU side of the reference's training step built so far (BASELINE configs[4]; reference:
network_functions.py:210-280 train_batch_2_loss, 168-208 calc_loss, c_network.py:187-226 forward in TRAIN mode, 243-261).

What runs on the GPU (fp32, sm_100a kernels only):
  forward   the whole train-mode C_NETWORK.forward — every ComplexBatchNorm2d with BATCH statistics and the running-stat
            update (dcs_cbn_train_fwd), un-folded convs, ComplexLSTM, fc, attentions, decoder[6], bound_cRM x2, combine, the
            three iSTFTs — and calc_loss (noise_loss_type 6 / speech_loss_type 0, the config.py defaults) from dcs_si_snr;
  backward  the first stage: loss -> waveform gradients (dcs_si_snr) -> iSTFT adjoint (dcs_istft_adjoint) -> mask-tail adjoint
            (dcs_mask_tail_bwd: polar, combine, bound_cRM x2) -> decoder[6] dgrad (the forward conv kernel with role-swapped
            weights) -> up-sampling / concat adjoint (dcs_upcat_adjoint) -> gradients w.r.t. decoder[5]'s attended output and
            skip[6]; plus the train-mode BatchNorm backward (dcs_cbn_train_bwd) as a stand-alone stage.
Still open (raises NotImplementedError): the rest of the backward chain (attention, LSTM BPTT, conv wgrad), the optimizer.
Dropout: the parity configuration sets both probabilities to 0; non-zero dropout is refused.
"""
import torch

from . import _lib as L
from . import ops, packing, train_ops as T
from .engine import KERNEL_E, STRIDE_E, UPSAMPLE, _sd_tensor_dict


class TrainStep:
    def __init__(self, model, variant="dcs", speech_alpha=0.7, atan2_eps=10e-7):
        assert variant in ("dcs", "dc")
        self.model, self.variant, self.alpha, self.eps = model, variant, float(speech_alpha), float(atan2_eps)
        hp = getattr(model, "hparams", {})
        if float(hp.get("dropout_conv", 0.0)) != 0.0 or float(hp.get("dropout_fc", 0.0)) != 0.0:
            raise NotImplementedError("dcsnet_b200.TrainStep: dropout is not built (set hparams dropout_conv = dropout_fc = 0)")
        self.L = int(hp.get("no_of_layers", 7))
        self._packed_key = None

    # ------------------------------------------------------------------ operands (un-folded: BN runs on batch statistics)
    def _pack(self, device):
        sd = _sd_tensor_dict(self.model)
        key = (str(device),) + tuple((v.data_ptr(), v._version) for v in sd.values())
        if key == self._packed_key:
            return
        Lr = self.L
        self.enc, self.dec = [], []
        for i in range(Lr):
            p = f"encoder.{i}.0."
            self.enc.append(packing.PackedConv(sd[p + "conv_r.weight"], sd[p + "conv_i.weight"], sd[p + "conv_r.bias"], sd[p + "conv_i.bias"],
                                               stride=STRIDE_E[i], act=L.ACT_NONE, device=device))
        for i in range(Lr):
            p = f"decoder.{i}." if i == Lr - 1 else f"decoder.{i}.0."
            self.dec.append(packing.PackedConv(sd[p + "conv_tran_r.weight"], sd[p + "conv_tran_i.weight"], sd[p + "conv_tran_r.bias"],
                                               sd[p + "conv_tran_i.bias"], transposed=True, up=UPSAMPLE[i], act=L.ACT_NONE, device=device))
        self.skip_ca = [packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", device) for i in range(Lr)]
        self.skip_sa = [packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", device) for i in range(Lr)]
        self.dec_ca = [packing.pack_channel_attention(sd, f"decoder_attention.{2 * i}.", device) for i in range(Lr - 1)]
        self.dec_sa = [packing.pack_spatial_attention(sd, f"decoder_attention.{2 * i + 1}.", device) for i in range(Lr - 1)]
        self.lstm = packing.pack_lstm(sd, "lstm.", device)
        self.fc = packing.PackedConv(sd["fc.fc_r.weight"][:, :, None, None], sd["fc.fc_i.weight"][:, :, None, None], sd["fc.fc_r.bias"],
                                     sd["fc.fc_i.bias"], device=device)
        p6 = f"decoder.{Lr - 1}."
        self.dec6_dgrad = T.dgrad_conv(sd[p6 + "conv_tran_r.weight"], sd[p6 + "conv_tran_i.weight"], transposed=True, device=device)
        self._packed_key = key

    def _bn(self, x, prefix, act):
        """Train-mode ComplexBatchNorm2d on the module's own parameters / buffers (running statistics updated in place)."""
        m = self.model.get_submodule(prefix)
        y, saved, _ = T.cbn_train_fwd(x, m.weight.detach(), m.bias.detach(), m.running_mean, m.running_covar, m.num_batches_tracked,
                                      act=act, eps=m.eps, momentum=m.momentum if m.momentum is not None else T.BN_MOMENTUM)
        return y, saved

    def _attention(self, x, ca, w7):
        B, H, W, Cn, _ = x.shape
        sums = ops.zero_(torch.empty(B, Cn, 2, dtype=torch.int64, device=x.device))
        ops.chan_pool(x, sums)
        gate = torch.empty(B, Cn, 2, dtype=torch.float32, device=x.device)
        stats = torch.empty(B, H * W, 4, dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        ops.spat_stats(x, None, stats, sums=sums, ca=ca, gate_out=gate)
        ops.spat_apply(x, gate, stats, w7, y)
        return y

    # ------------------------------------------------------------------ forward + loss
    def forward(self, noise_spec, noisy_spec, clean_spec):
        """train_batch_2_loss (network_functions.py:210-280) on (B, 256, T) complex64 CUDA spectrograms.  Returns the dict of
        losses (device scalars) and keeps what the backward needs in self.saved."""
        L.require_cuda(noise_spec, noisy_spec, clean_spec)
        dev = noisy_spec.device
        with torch.cuda.device(dev):
            return self._forward(noise_spec.contiguous(), noisy_spec.contiguous(), clean_spec.contiguous(), dev)

    def _forward(self, noise_spec, Y, clean_spec, dev):
        self._pack(dev)
        Lr = self.L
        B, F, Tn = Y.shape
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)   # noqa: E731
        sv = {}
        x, sv["bn0"] = self._bn(torch.view_as_real(Y).view(B, F, Tn, 1, 2), "initial_batchnorm", L.ACT_NONE)
        enc = []
        H, W = F, Tn
        for i in range(Lr):
            H, W = ops.conv_out_hw(self.enc[i], H, W)
            pre = ops.cconv(self.enc[i], x, None, new(B, H, W, self.enc[i].cout, 2))
            x, sv[f"enc{i}"] = self._bn(pre, f"encoder.{i}.1", L.ACT_RELU)
            sv[f"enc{i}_pre"] = pre
            enc.append(x)
        S = H * W
        lat = new(B, H, W, 128, 2)
        ws = torch.empty(ops.clstm_workspace_bytes(B, S) // 4, dtype=torch.float32, device=dev)
        ops.clstm(x.view(B, S, x.shape[3], 2), lat.view(B, S, 128, 2), self.lstm, ws, use_tc=False)
        d = ops.cconv(self.fc, lat.view(B, 1, S, 128, 2), None, new(B, 1, S, self.fc.cout, 2)).view(B, H, W, self.fc.cout, 2)
        for i in range(Lr):
            skip = self._attention(enc[Lr - 1 - i], self.skip_ca[i], self.skip_sa[i])
            H, W = H * UPSAMPLE[i][0], W * UPSAMPLE[i][1]
            pre = ops.cconv(self.dec[i], d, skip, new(B, H, W, self.dec[i].cout, 2))
            if i == Lr - 1:
                sv["d5"], sv["skip6"] = d, skip
                raw = torch.view_as_complex(pre.view(B, H, W, 2))
                break
            a, sv[f"dec{i}"] = self._bn(pre, f"decoder.{i}.1", L.ACT_LRELU)
            d = self._attention(a, self.dec_ca[i], self.dec_sa[i])
        # ---- mask tail + the three waveforms + calc_loss
        est_clean = torch.empty_like(Y)
        est_noise = torch.empty_like(Y) if self.variant == "dcs" else None
        ops.mask_combine(raw, Y, est_clean, noise_spec=est_noise, atan2_eps=self.eps,
                         combine=L.COMBINE_DCS if self.variant == "dcs" else L.COMBINE_DC, exact_polar=True)
        wave = lambda s: ops.istft(s, atan2_eps=self.eps, exact_polar=True)   # noqa: E731
        clean_audio, est_clean_audio = wave(clean_spec), wave(est_clean)
        out = {}
        si_c, _ = T.si_snr(clean_audio, est_clean_audio)
        out["speech_loss"] = self.alpha * (-si_c.mean())
        if self.variant == "dcs":
            noise_audio, est_noise_audio = wave(noise_spec), wave(est_noise)
            si_n, _ = T.si_snr(noise_audio, est_noise_audio)
            out["noise_loss"] = 1 - self.alpha * (-si_n.mean())        # network_functions.py:195-196, precedence as written
            out["train_loss"] = out["noise_loss"] + out["speech_loss"]
            sv.update(noise_audio=noise_audio, est_noise_audio=est_noise_audio)
        else:
            out["noise_loss"], out["train_loss"] = None, out["speech_loss"]
        sv.update(raw=raw, Y=Y, clean_audio=clean_audio, est_clean_audio=est_clean_audio, T=Tn)
        self.saved = sv
        return out

    # ------------------------------------------------------------------ backward, first stage
    def backward_first_stage(self):
        """d train_loss / d (decoder[5] attended output, skip[6]) and the intermediate gradients, from the saved forward.
        Returns dict(g_clean_wave, g_noise_wave, d_raw, g_d5, g_skip6)."""
        sv = self.saved
        dev = sv["raw"].device
        with torch.cuda.device(dev):
            # total = [1 - alpha (-SiSNR_n)] + alpha (-SiSNR_c)  =>  d/d s_hat = -alpha dSiSNR_c, d/d n_hat = +alpha dSiSNR_n
            _, g_clean = T.si_snr(sv["clean_audio"], sv["est_clean_audio"], grad_scale=-self.alpha)
            g_noise = None
            if self.variant == "dcs":
                _, g_noise = T.si_snr(sv["noise_audio"], sv["est_noise_audio"], grad_scale=self.alpha)
            gS = T.istft_adjoint(g_clean, sv["T"])
            gN = T.istft_adjoint(g_noise, sv["T"]) if g_noise is not None else None
            d_raw = T.mask_tail_bwd(sv["raw"].contiguous(), sv["Y"], gS, gN, self.eps)
            B, F, Tn = d_raw.shape
            dy = torch.view_as_real(d_raw).view(B, F, Tn, 1, 2)
            c0, c1 = sv["d5"].shape[3], sv["skip6"].shape[3]
            g_up = ops.cconv(self.dec6_dgrad, dy, None, torch.empty(B, F, Tn, c0 + c1, 2, dtype=torch.float32, device=dev))
            g_d5, g_skip6 = T.upcat_adjoint(g_up, c0, c1, UPSAMPLE[self.L - 1])
        return dict(g_clean_wave=g_clean, g_noise_wave=g_noise, d_raw=d_raw, g_d5=g_d5, g_skip6=g_skip6)

    def backward(self):
        raise NotImplementedError("dcsnet_b200.TrainStep: only the first backward stage is built (backward_first_stage); attention / "
                                  "LSTM BPTT / conv wgrad kernels and the optimizer are SURVEY 8f rank 2 work still open")
