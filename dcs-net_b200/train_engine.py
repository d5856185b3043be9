"""TrainStep — the GPU side of the reference's training step (BASELINE configs[4]; reference:
network_functions.py:210-280 train_batch_2_loss, 168-208 calc_loss, c_network.py:187-226 forward in TRAIN mode, 229-235 Adam-amsgrad,
243-261 training_step; config.py:48-49 gradient clip).

Everything runs as sm_100a kernels of libdcsnet_sm100a.so (fp32; no autograd, no PyTorch-op fallback):
  forward   the whole train-mode C_NETWORK.forward — every ComplexBatchNorm2d with BATCH statistics and the running-stat
            update (dcs_cbn_train_fwd), un-folded convs, ComplexLSTM in training form (dcs_sgemm projections + dcs_lstm_train_fwd,
            gates / cells saved), fc, attentions (saved gates / statistics), dropout (dcs_dropout, Philox), decoder[6],
            bound_cRM x2, combine, the three iSTFTs — and calc_loss (noise_loss_type 6 / speech_loss_type 0) from dcs_si_snr;
  backward  loss -> waveform gradients -> iSTFT adjoint -> mask-tail adjoint -> per decoder stage [dropout^T, attention backward,
            LeakyReLU mask, train-mode BN backward, ComplexConvTranspose2d wgrad (dcs_upcat_fwd + dcs_wgrad) / bias / dgrad (the
            forward conv kernel with role-swapped weights) + up-sampling / concat adjoint, skip-attention backward] -> fc ->
            ComplexLSTM BPTT (dcs_lstm_train_bwd + GEMM-side gradients) -> per encoder layer [ReLU mask, BN backward, wgrad, bias,
            strided dgrad = dcs_dilate + forward conv kernel] -> initial_batchnorm; gradients land in the parameters' .grad;
  step      global-norm clip + Adam-amsgrad with L2 weight decay on flat fp32 buffers (dcs_sumsq, dcs_adam_amsgrad).
Checked against tests/golden/train_step.pt (the reference's own train_batch_2_loss + backward()): losses, running statistics, all
198 parameter gradients, the global gradient norm (tests/test_train_gpu.py).
"""
import torch

from . import _lib as L
from . import ops, packing, train_ops as T
from .engine import KERNEL_E, STRIDE_E, UPSAMPLE, _sd_tensor_dict

ADAM_BETAS = (0.9, 0.999)        # torch.optim.Adam defaults (c_network.py:229-235 passes lr, eps, weight_decay, amsgrad only)


class TrainStep:
    def __init__(self, model, variant="dcs", speech_alpha=None, atan2_eps=None, seed=0, mode="fp32"):
        """mode: "fp32" = CUDA-core FFMA convolutions (the parity mode); "tf32" = the forward and data-gradient convolutions on the
        tensor cores (tcgen05 kind::tf32 through dcs_cconv2d_tc_fwd, fp32 storage, fp32 accumulation)."""
        assert variant in ("dcs", "dc") and mode in ("fp32", "tf32", "bf16")
        self.mode = mode
        # "bf16": the saved activations are stored in bf16 and the forward convolutions run in kind::f16 on them; gradients, BatchNorm
        # statistics, attention gates, the LSTM and the master weights stay fp32 (the data-gradient convolutions read fp32 gradients: tf32)
        self.act_dtype = torch.bfloat16 if mode == "bf16" else torch.float32
        hp = getattr(model, "hparams", {})
        self.model, self.variant, self.hp = model, variant, hp
        self.alpha = float(speech_alpha if speech_alpha is not None else hp.get("speech_alpha", 0.7))
        self.eps = float(atan2_eps if atan2_eps is not None else hp.get("atan2_eps", 10e-7))
        # the reference builds torch.nn.Dropout(p) from dropout_conv / dropout_fc alone (c_network.py:166-167)
        self.p_conv, self.p_fc = float(hp.get("dropout_conv", 0.0)), float(hp.get("dropout_fc", 0.0))
        self.seed, self.steps_done = int(seed), 0
        self.L = int(hp.get("no_of_layers", 7))
        self._packed_key = None
        self._stale = False
        self.gather = None
        self.opt = None
        self.saved = None

    # ------------------------------------------------------------------ operands (un-folded: BN runs on batch statistics)
    def _pack(self, device):
        sd = _sd_tensor_dict(self.model)
        key = (str(device),) + tuple((v.data_ptr(), v._version) for v in sd.values())
        if key == self._packed_key and not self._stale:
            return
        if self.opt is not None and str(self.flat_param.device) != str(device):
            raise RuntimeError("TrainStep: the model moved to another device after init_optimizer(); build a new TrainStep for it "
                               "(the flat parameter / gradient / Adam-state buffers and the operand tables live on the first device)")
        if self.gather is not None and self._packed_key is not None and key[0] == self._packed_key[0]:
            self.gather.run()                 # one launch: every operand rebuilt from the flat parameter buffer
            self._packed_key, self._stale = key, False
            return
        Lr = self.L
        tf = self.mode in ("tf32", "bf16")
        tcd = torch.bfloat16 if self.mode == "bf16" else None
        self.enc, self.dec, self.enc_dgrad, self.dec_dgrad = [], [], [], []
        for i in range(Lr):
            p = f"encoder.{i}.0."
            self.enc.append(packing.PackedConv(sd[p + "conv_r.weight"], sd[p + "conv_i.weight"], sd[p + "conv_r.bias"], sd[p + "conv_i.bias"],
                                               stride=STRIDE_E[i], act=L.ACT_NONE, device=device, want_tf32=tf and tcd is None, tc_dtype=tcd))
            # data gradient of the strided conv as a sub-pixel phase convolution of the un-dilated gradient (train_ops.PhasePack)
            self.enc_dgrad.append(T.PhasePack(sd[p + "conv_r.weight"], sd[p + "conv_i.weight"], STRIDE_E[i], device, want_tf32=tf))
        for i in range(Lr):
            p = f"decoder.{i}." if i == Lr - 1 else f"decoder.{i}.0."
            self.dec.append(packing.PackedConv(sd[p + "conv_tran_r.weight"], sd[p + "conv_tran_i.weight"], sd[p + "conv_tran_r.bias"],
                                               sd[p + "conv_tran_i.bias"], transposed=True, up=UPSAMPLE[i], act=L.ACT_NONE, device=device, want_tf32=tf and tcd is None,
                                               tc_dtype=tcd))
            # data gradient per source (decoder path d: the first half of the input channels, skip: the second), so that each GEMM's
            # N = 2 * channels stays within the tensor-core kernel's 256 columns
            c0 = sd[p + "conv_tran_r.weight"].shape[0] // 2
            self.dec_dgrad.append(tuple(T.dgrad_conv(sd[p + "conv_tran_r.weight"][sl], sd[p + "conv_tran_i.weight"][sl], transposed=True, device=device,
                                                     want_tf32=tf) for sl in (slice(0, c0), slice(c0, None))))
        self.skip_ca = [packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", device) for i in range(Lr)]
        self.skip_sa = [packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", device) for i in range(Lr)]
        self.dec_ca = [packing.pack_channel_attention(sd, f"decoder_attention.{2 * i}.", device) for i in range(Lr - 1)]
        self.dec_sa = [packing.pack_spatial_attention(sd, f"decoder_attention.{2 * i + 1}.", device) for i in range(Lr - 1)]
        # ComplexLSTM (c_network.py:12-51): per layer l and lstm j (real_lstm, imag_lstm): W_ih of both directions stacked (512, D),
        # the summed biases (512,), W_hh (2 j, 2 dirs, 256, 64)
        names, sfx = ("real_lstm", "imag_lstm"), ("", "_reverse")
        f = lambda k: sd["lstm." + k].to(device=device, dtype=torch.float32)   # noqa: E731
        self.lstm_wih = [[torch.cat([f(f"{n}.weight_ih_l{l}{s}") for s in sfx], 0).contiguous() for n in names] for l in range(2)]
        self.lstm_b = [[torch.cat([f(f"{n}.bias_ih_l{l}{s}") + f(f"{n}.bias_hh_l{l}{s}") for s in sfx], 0).contiguous() for n in names] for l in range(2)]
        self.lstm_whh = [torch.stack([torch.stack([f(f"{n}.weight_hh_l{l}{s}") for s in sfx], 0) for n in names], 0).contiguous() for l in range(2)]
        self.fc = packing.PackedConv(sd["fc.fc_r.weight"][:, :, None, None], sd["fc.fc_i.weight"][:, :, None, None], sd["fc.fc_r.bias"],
                                     sd["fc.fc_i.bias"], device=device, want_tf32=tf)
        self.fc_dgrad = T.dgrad_conv(sd["fc.fc_r.weight"][:, :, None, None], sd["fc.fc_i.weight"][:, :, None, None], transposed=False, device=device,
                                     want_tf32=tf)
        self._packed_key, self._stale = key, False

    def _conv(self, pk, src0, src1, dst):
        """One complex convolution: tcgen05 kind::tf32 when the mode and the layer allow it (>= 4 complex channels per source = 32-byte
        rows for the operand loader, 2 cout <= 256), CUDA-core FFMA otherwise."""
        c1 = 0 if src1 is None else src1.shape[3]
        have = pk.w_tc is not None if src0.dtype in ops.H16 else pk.w_tc32 is not None
        tc = self.mode != "fp32" and have and src0.shape[3] % 4 == 0 and c1 % 4 == 0 and 2 * pk.cout <= 256
        if dst.dtype in ops.H16 and src0.dtype == torch.float32 and not tc:
            raise NotImplementedError("TrainStep: an fp32 -> 16-bit convolution needs the tensor-core path")
        return ops.cconv(pk, src0, src1, dst, use_tc=tc)

    def _bn(self, x, prefix, act, out=None):
        """Train-mode ComplexBatchNorm2d on the module's own parameters / buffers (running statistics updated in place)."""
        m = self.model.get_submodule(prefix)
        y, saved, _ = T.cbn_train_fwd(x, m.weight.detach(), m.bias.detach(), m.running_mean, m.running_covar, m.num_batches_tracked,
                                      act=act, y=out, eps=m.eps, momentum=m.momentum if m.momentum is not None else T.BN_MOMENTUM)
        return y, saved

    def _drop(self, x, p, tag):
        """torch.nn.Dropout(p) on view_as_real (c_network.py:195-196 / 203-204 / 221-222); the (seed, offset) pair is kept for the backward."""
        if p <= 0.0:
            return x
        off = self._drop_off
        self._drop_off += (x.numel() + 3) // 4
        self.saved["drop_" + tag] = (p, off)
        return T.dropout(x, p, self.seed + 1000003 * self.steps_done, off)

    def _drop_bwd(self, g, tag):
        rec = self.saved.get("drop_" + tag)
        return g if rec is None else T.dropout(g.contiguous(), rec[0], self.seed + 1000003 * self.steps_done, rec[1])

    # ------------------------------------------------------------------ forward + loss
    def forward(self, noise_spec, noisy_spec, clean_spec):
        """train_batch_2_loss (network_functions.py:210-280) on (B, 256, T) complex64 CUDA spectrograms.  Returns the dict of
        losses (device scalars) and keeps what the backward needs in self.saved."""
        L.require_cuda(noise_spec, noisy_spec, clean_spec)
        dev = noisy_spec.device
        with torch.cuda.device(dev):
            return self._forward(noise_spec.contiguous(), noisy_spec.contiguous(), clean_spec.contiguous(), dev)

    def _lstm_forward(self, x, sv):
        """ComplexLSTM in training form: x (B, S, 128, 2) -> (B, S, 128, 2); four real LSTM passes (R / I on re / im) as two weight
        groups x two part planes = 4 B sequences per direction."""
        B, S, D, _ = x.shape
        X0 = T.cplx_split(x.view(B * S, D, 2))                               # (2 part, B S, D)
        inp = [X0.view(2 * B * S, D)] * 2                                      # layer-0 input of lstm j (shared)
        sv["lstm_in"], sv["lstm"] = [inp], []
        for l in range(2):
            pre = torch.empty(2, 2 * B * S, 512, dtype=torch.float32, device=x.device)        # (j, (p, b, s), (dir, 4H))
            for j in range(2):
                T.sgemm(inp[j], self.lstm_wih[l][j], self.lstm_b[l][j], out=pre[j])
            h, gates, cells = T.lstm_train_fwd(pre.view(4 * B, S, 2, 256), self.lstm_whh[l], 2)  # h (4B = (j, p, b), S, 2, 64)
            sv["lstm"].append(dict(h=h, gates=gates, cells=cells))
            inp = [h.view(2, 2 * B * S, 128)[j] for j in range(2)]
            sv["lstm_in"].append(inp)
        return T.clstm_combine(h.view(2, 2, B * S * 128)).view(B, S, 128, 2)

    def _forward(self, noise_spec, Y, clean_spec, dev):
        self._pack(dev)
        Lr = self.L
        B, F, Tn = Y.shape
        new = lambda *s: torch.empty(*s, dtype=self.act_dtype, device=dev)   # noqa: E731  (activations: fp32, or bf16 storage)
        sv = self.saved = {}
        self._drop_off = 0
        x0 = torch.view_as_real(Y).view(B, F, Tn, 1, 2)
        x, sv["bn0"] = self._bn(x0, "initial_batchnorm", L.ACT_NONE, out=new(B, F, Tn, 1, 2))
        sv["x0"] = x0
        enc = [x]                                                             # enc[i] = input of encoder i; enc[i + 1] its (dropped) output
        H, W = F, Tn
        for i in range(Lr):
            H, W = ops.conv_out_hw(self.enc[i], H, W)
            pre = self._conv(self.enc[i], x, None, new(B, H, W, self.enc[i].cout, 2))
            x, sv[f"enc{i}"] = self._bn(pre, f"encoder.{i}.1", L.ACT_RELU)
            sv[f"enc{i}_pre"] = pre
            x = self._drop(x, self.p_conv, f"enc{i}")
            enc.append(x)
        sv["enc"] = enc
        S = H * W
        lat = self._lstm_forward(x.view(B, S, x.shape[3], 2), sv)
        sv["lat"] = lat
        d = self._conv(self.fc, lat.view(B, 1, S, 128, 2), None, new(B, 1, S, self.fc.cout, 2)).view(B, H, W, self.fc.cout, 2)
        d = self._drop(d, self.p_fc, "fc")
        sv["dec_in"], sv["skip"], sv["skip_att"], sv["dec_att"], sv["dec_act"] = [], [], [], [], []
        for i in range(Lr):
            skip, att = T.attention_fwd_saved(enc[Lr - i], self.skip_ca[i], self.skip_sa[i])
            sv["dec_in"].append(d), sv["skip"].append(skip), sv["skip_att"].append(att)
            H, W = H * UPSAMPLE[i][0], W * UPSAMPLE[i][1]
            last = i == Lr - 1        # decoder[6] writes the raw mask in fp32 (complex64 view)
            pre = self._conv(self.dec[i], d, skip, torch.empty(B, H, W, self.dec[i].cout, 2, dtype=torch.float32 if last else self.act_dtype, device=dev))
            if i == Lr - 1:
                pre = self._drop(pre, self.p_conv, f"dec{i}")
                sv["d5"], sv["skip6"] = d, skip
                raw = torch.view_as_complex(pre.view(B, H, W, 2))
                break
            a, sv[f"dec{i}"] = self._bn(pre, f"decoder.{i}.1", L.ACT_LRELU)
            sv[f"dec{i}_pre"] = pre
            d, att = T.attention_fwd_saved(a, self.dec_ca[i], self.dec_sa[i])
            sv["dec_att"].append(att)
            d = self._drop(d, self.p_conv, f"dec{i}")
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
        return out

    # ------------------------------------------------------------------ backward
    def _tail_backward(self):
        """loss -> waveform gradients -> iSTFT adjoint -> mask-tail adjoint: d train_loss / d raw (after decoder[6]'s dropout)."""
        sv = self.saved
        # total = [1 - alpha (-SiSNR_n)] + alpha (-SiSNR_c)  =>  d/d s_hat = -alpha dSiSNR_c, d/d n_hat = +alpha dSiSNR_n
        _, g_clean = T.si_snr(sv["clean_audio"], sv["est_clean_audio"], grad_scale=-self.alpha)
        g_noise = None
        if self.variant == "dcs":
            _, g_noise = T.si_snr(sv["noise_audio"], sv["est_noise_audio"], grad_scale=self.alpha)
        gS = T.istft_adjoint(g_clean, sv["T"])
        gN = T.istft_adjoint(g_noise, sv["T"]) if g_noise is not None else None
        d_raw = T.mask_tail_bwd(sv["raw"].contiguous(), sv["Y"], gS, gN, self.eps)
        return g_clean, g_noise, d_raw

    def backward_first_stage(self):
        """d train_loss / d (decoder[5] attended output, skip[6]) and the intermediate gradients, from the saved forward.
        Returns dict(g_clean_wave, g_noise_wave, d_raw, g_d5, g_skip6)."""
        sv = self.saved
        dev = sv["raw"].device
        with torch.cuda.device(dev):
            g_clean, g_noise, d_raw = self._tail_backward()
            B, F, Tn = d_raw.shape
            dy = self._drop_bwd(torch.view_as_real(d_raw).view(B, F, Tn, 1, 2), f"dec{self.L - 1}")
            g_d5, g_skip6 = self._dec_dgrad(self.L - 1, dy.contiguous(), sv["d5"].shape[3], sv["skip6"].shape[3])
        return dict(g_clean_wave=g_clean, g_noise_wave=g_noise, d_raw=d_raw, g_d5=g_d5, g_skip6=g_skip6)

    def _dec_dgrad(self, i, dpre, c0, c1):
        """Data gradient of decoder stage i's cat + up-sampling + ComplexConvTranspose2d: per source, the forward conv kernel with the
        role-swapped weights, then the up-sampling adjoint (sum over each up_h x up_w block)."""
        B, HH, WW = dpre.shape[0], dpre.shape[1], dpre.shape[2]
        out = []
        for pk, c in zip(self.dec_dgrad[i], (c0, c1)):
            g_up = self._conv(pk, dpre, None, torch.empty(B, HH, WW, c, 2, dtype=torch.float32, device=dpre.device))
            out.append(T.upcat_adjoint(g_up, c, 0, UPSAMPLE[i])[0])
        return out

    def _grad(self, name):
        """The parameter's .grad tensor (allocated on first use; GradBuckets / FlatAdam make it a view into a flat buffer)."""
        p = self._params[name]
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        return p.grad

    def _att_grads(self, chan_prefix, spat_prefix):
        g = self._grad
        return dict(dw1_r=g(chan_prefix + "fc.0.conv_r.weight"), dw1_i=g(chan_prefix + "fc.0.conv_i.weight"),
                    dw2_r=g(chan_prefix + "fc.2.conv_r.weight"), dw2_i=g(chan_prefix + "fc.2.conv_i.weight"),
                    dw7_r=g(spat_prefix + "conv1.conv_r.weight"), dw7_i=g(spat_prefix + "conv1.conv_i.weight"))

    def _bn_bwd(self, x, dz, prefix, key, conv_bias=None):
        """Train-mode BN backward; conv_bias = (prefix, names) of the convolution in front: its two bias gradients (per-channel sums of
        dx) come out of the same pass."""
        m = self.model.get_submodule(prefix)
        cb = None if conv_bias is None else (self._grad(conv_bias[0] + conv_bias[1][0] + ".bias"), self._grad(conv_bias[0] + conv_bias[1][1] + ".bias"))
        dx, dw, db = T.cbn_train_bwd(x, dz, self.saved[key], m.weight.detach(), conv_bias_grads=cb)
        self._grad(prefix + ".weight").copy_(dw)
        self._grad(prefix + ".bias").copy_(db)
        return dx

    def _wgrad_on_tc(self, cin, cout):
        """tf32 mode: weight gradients of every layer with >= 8 complex channels on both sides run on tcgen05 (dcs_wgrad_tc16) from bf16
        copies of the two operands (fp32 accumulation over the pixels); encoder[0] / decoder[6] stay on the few-channel CUDA-core kernel."""
        return self.mode in ("tf32", "bf16") and cin >= 8 and cout >= 8

    def _conv_param_grads(self, x, dpre, prefix, names, kernel, stride, transposed, bias=True):
        if self._wgrad_on_tc(x.shape[3], dpre.shape[3]):
            x = x if x.dtype == torch.bfloat16 else (T.to_h16(x) if x.dtype == torch.float32 else T.to_h16(T.to_f32(x)))
            dy16 = T.to_h16(dpre)
        else:
            x = x if x.dtype == torch.float32 else T.to_f32(x)       # few-channel layers: the CUDA-core fp32 kernel
            dy16 = dpre
        T.cwgrad_generic(x, dy16, kernel, stride, transposed=transposed, dw_r=self._grad(prefix + names[0] + ".weight"),
                         dw_i=self._grad(prefix + names[1] + ".weight"))
        if bias:
            T.colsum(dpre.view(-1, 2 * dpre.shape[-2]), mode=1, out0=self._grad(prefix + names[0] + ".bias"), out1=self._grad(prefix + names[1] + ".bias"))

    def _lstm_backward(self, g_lat):
        """g_lat (B, S, 128, 2): gradient of the ComplexLSTM output -> gradient of its input (B, S, 128, 2); parameter gradients of
        real_lstm / imag_lstm (both layers, both directions)."""
        sv = self.saved
        B, S = g_lat.shape[0], g_lat.shape[1]
        names, sfx = ("real_lstm", "imag_lstm"), ("", "_reverse")
        dH = T.clstm_combine_bwd(g_lat.contiguous().view(B * S * 128, 2)).view(4 * B, S, 2, 64)
        dX = None
        for l in (1, 0):
            st = sv["lstm"][l]
            dpre = T.lstm_train_bwd(self.lstm_whh[l], st["gates"], st["cells"], dH, 2)          # (4B, S, 2, 256)
            dpre_j = dpre.view(2, 2 * B * S, 512)
            h_j = st["h"].view(2, 2 * B, 1, S, 128)
            D = self.lstm_wih[l][0].shape[1]
            d_in = torch.empty(2, 2 * B * S, D, dtype=torch.float32, device=g_lat.device) if l == 1 else None
            for j in range(2):
                x_in = sv["lstm_in"][l][j]                                                       # (2 B S, D)
                dwp = T.wgrad(x_in.view(2 * B, 1, S, D), dpre_j[j].view(2 * B, 1, S, 512), [(0, 0)])[0]   # (D, 512)
                for d in range(2):
                    pn = f"lstm.{names[j]}."
                    T.transpose_into(dwp[:, d * 256:(d + 1) * 256], self._grad(pn + f"weight_ih_l{l}{sfx[d]}"))
                    T.colsum(dpre_j[j][:, d * 256:(d + 1) * 256], mode=0, out0=self._grad(pn + f"bias_ih_l{l}{sfx[d]}"),
                             out1=self._grad(pn + f"bias_hh_l{l}{sfx[d]}"))
                    dwh = T.wgrad(h_j[j][..., d * 64:(d + 1) * 64], dpre_j[j].view(2 * B, 1, S, 512)[..., d * 256:(d + 1) * 256],
                                  [(0, 1 if d else -1)])[0]                                     # (64, 256)
                    T.transpose_into(dwh, self._grad(pn + f"weight_hh_l{l}{sfx[d]}"))
                if l == 1:
                    T.sgemm(dpre_j[j], self.lstm_wih[l][j], out=d_in[j], b_is_nk=False)
                else:
                    dX = T.sgemm(dpre_j[j], self.lstm_wih[l][j], out=dX, b_is_nk=False, accumulate=j > 0)
            if l == 1:
                dH = d_in.view(4 * B, S, 2, 64)
        return T.cplx_merge(dX.view(2, B * S * 128)).view(B, S, 128, 2)

    def backward(self):
        """The whole backward pass of the last forward(): fills .grad of every parameter (198 tensors).  Returns a dict of a few
        intermediate gradients for inspection."""
        sv = self.saved
        if sv is None:
            raise RuntimeError("TrainStep.backward: call forward() first")
        dev = sv["raw"].device
        self._params = dict(self.model.named_parameters())
        with torch.cuda.device(dev):
            return self._backward(dev)

    def _backward(self, dev):
        sv, Lr = self.saved, self.L
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)   # noqa: E731
        _, _, d_raw = self._tail_backward()
        B, F, Tn = d_raw.shape
        enc = sv["enc"]
        skip_grad = [None] * (Lr + 1)          # skip_grad[k] = (dx, chan_const): the skip path's gradient into enc[k]
        g = torch.view_as_real(d_raw).view(B, F, Tn, 1, 2)
        info = {}
        for i in range(Lr - 1, -1, -1):
            g = self._drop_bwd(g, f"dec{i}")
            if i == Lr - 1:
                dpre, prefix = g.contiguous(), f"decoder.{i}."
            else:
                att = sv["dec_att"][i]
                dxa, cc, _ = T.attention_bwd(att["x"], g.contiguous(), att["gate_c"], att["stats"], att["gate_s"], att["sums"], self.dec_ca[i],
                                             self.dec_sa[i], grads=self._att_grads(f"decoder_attention.{2 * i}.", f"decoder_attention.{2 * i + 1}."))
                dz = T.act_bwd(att["x"], dxa, L.ACT_LRELU, None, cc)
                prefix = f"decoder.{i}.0."
                dpre = self._bn_bwd(sv[f"dec{i}_pre"], dz, f"decoder.{i}.1", f"dec{i}", conv_bias=(prefix, ("conv_tran_r", "conv_tran_i")))
            d_in, skip = sv["dec_in"][i], sv["skip"][i]
            if i == Lr - 1 and dpre.shape[3] == 1 and UPSAMPLE[i] == (2, 2) and d_in.shape[3] + skip.shape[3] <= 16:
                # decoder[6]: one output channel -> data, weight and bias gradients in ONE kernel (no up-sampled input, no full-resolution dgrad)
                g, g_skip = T.dec6_bwd(d_in, skip, dpre, self._params[prefix + "conv_tran_r.weight"].detach(), self._params[prefix + "conv_tran_i.weight"].detach(),
                                       self._grad(prefix + "conv_tran_r.weight"), self._grad(prefix + "conv_tran_i.weight"),
                                       self._grad(prefix + "conv_tran_r.bias"), self._grad(prefix + "conv_tran_i.bias"))
                info["g_d5"], info["g_skip6"] = g, g_skip
                att = sv["skip_att"][i]
                dxs, ccs, _ = T.attention_bwd(att["x"], g_skip, att["gate_c"], att["stats"], att["gate_s"], att["sums"], self.skip_ca[i], self.skip_sa[i],
                                              grads=self._att_grads(f"skip_attention.{2 * i}.", f"skip_attention.{2 * i + 1}."))
                skip_grad[Lr - i] = (dxs, ccs)
                continue
            on_tc = self._wgrad_on_tc(d_in.shape[3] + skip.shape[3], dpre.shape[3])
            z = T.upcat_fwd(d_in, skip, UPSAMPLE[i], dtype=torch.bfloat16 if on_tc else torch.float32)
            self._conv_param_grads(z, dpre, prefix, ("conv_tran_r", "conv_tran_i"), 3, (1, 1), True, bias=i == Lr - 1)
            del z
            g, g_skip = self._dec_dgrad(i, dpre, d_in.shape[3], skip.shape[3])
            if i == Lr - 1:
                info["g_d5"], info["g_skip6"] = g, g_skip
            att = sv["skip_att"][i]
            dxs, ccs, _ = T.attention_bwd(att["x"], g_skip, att["gate_c"], att["stats"], att["gate_s"], att["sums"], self.skip_ca[i], self.skip_sa[i],
                                          grads=self._att_grads(f"skip_attention.{2 * i}.", f"skip_attention.{2 * i + 1}."))
            skip_grad[Lr - i] = (dxs, ccs)
        # ---- fc (ComplexLinear as a 1x1 conv over the B*S latent rows) and the ComplexLSTM
        Hl, Wl = g.shape[1], g.shape[2]
        S = Hl * Wl
        g = self._drop_bwd(g, "fc").contiguous().view(B, 1, S, 128, 2)
        lat = sv["lat"].view(B, 1, S, 128, 2)
        self._conv_param_grads(lat, g, "fc.", ("fc_r", "fc_i"), 1, (1, 1), False)
        g_lat = self._conv(self.fc_dgrad, g, None, new(B, 1, S, 128, 2))
        g = self._lstm_backward(g_lat.view(B, S, 128, 2)).view(B, Hl, Wl, 128, 2)
        info["g_latent_in"] = g
        # ---- encoder, last layer first: enc[i + 1] feeds encoder i + 1 (or the LSTM) AND skip attention Lr - 1 - i
        for i in range(Lr - 1, -1, -1):
            dxs, ccs = skip_grad[i + 1]
            y = enc[i + 1]
            if sv.get(f"drop_enc{i}") is None:
                dz = T.act_bwd(y, g.contiguous(), L.ACT_RELU, dxs, ccs)
            else:
                tot = self._drop_bwd(T.act_bwd(None, g.contiguous(), L.ACT_NONE, dxs, ccs), f"enc{i}")
                dz = T.act_bwd(y, tot, L.ACT_RELU)
            dpre = self._bn_bwd(sv[f"enc{i}_pre"], dz, f"encoder.{i}.1", f"enc{i}", conv_bias=(f"encoder.{i}.0.", ("conv_r", "conv_i")))
            x_in = enc[i]
            self._conv_param_grads(x_in, dpre, f"encoder.{i}.0.", ("conv_r", "conv_i"), KERNEL_E[i], STRIDE_E[i], False, bias=False)
            Hi, Wi = x_in.shape[1], x_in.shape[2]
            if x_in.shape[3] == 1:       # encoder[0]: one input channel -> the direct gather kernel on the raw weights
                g = T.cconv_dgrad_cin1(dpre, self._params[f"encoder.{i}.0.conv_r.weight"].detach(), self._params[f"encoder.{i}.0.conv_i.weight"].detach(),
                                       Hi, Wi, STRIDE_E[i])
            else:
                if (Hi, Wi) != (dpre.shape[1] * STRIDE_E[i][0], dpre.shape[2] * STRIDE_E[i][1]):
                    raise NotImplementedError("TrainStep: encoder input sizes must be multiples of the strides (F % 128 == 0, T % 8 == 0, as "
                                              "the reference's decoder concatenation requires, c_network.py:214)")
                g = self._conv(self.enc_dgrad[i], dpre, None, new(B, Hi, Wi, x_in.shape[3], 2))
        self._bn_bwd(sv["x0"].contiguous(), g, "initial_batchnorm", "bn0")
        return info

    # ------------------------------------------------------------------ optimizer
    def init_optimizer(self, world_size=1):
        """Flat fp32 parameter / gradient / Adam-state buffers in GradBuckets order (decoder first); parameters and their .grad become
        views, so the backward kernels write the flat gradient directly and one fused kernel updates everything."""
        from .grad_sync import GradBuckets
        named = list(self.model.named_parameters())
        dev = named[0][1].device
        self.buckets = GradBuckets(named, device=dev, flat=True)
        n = self.buckets.numel
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for name, p in self.buckets.order:
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        self.opt = dict(m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev), vmax=torch.zeros(n, device=dev), step=0,
                        sumsq=torch.zeros(1, dtype=torch.float64, device=dev), ws=torch.empty(8 * 4096, dtype=torch.uint8, device=dev),
                        world=world_size)
        self._packed_key, self.gather = None, None
        self._pack(dev)                                   # host packing once (shapes, tap tables); the tensors are redirected below
        self._build_gather(dev)
        return self

    def _build_gather(self, dev):
        """(index, sign) tables of every packed operand w.r.t. the flat parameter buffer (train_pack.py): after an optimizer step ONE
        dcs_gather_pack launch per operand type rebuilds them — no host packing in the step."""
        from . import train_pack as TP
        off, o = {}, 0
        for name, p in self.buckets.order:
            off[name] = (o, tuple(p.shape))
            o += p.numel()
        leaf = lambda name, shape=None: TP.Sym.leaf(off[name][0], shape or off[name][1])   # noqa: E731
        gp = TP.GatherPack(self.flat_param)
        tf, Lr = self.mode in ("tf32", "bf16"), self.L
        tc16 = self.mode == "bf16"
        for i in range(Lr):
            p = f"encoder.{i}.0."
            wr, wi = leaf(p + "conv_r.weight"), leaf(p + "conv_i.weight")
            gp.add_packed_conv(self.enc[i], TP.sym_conv(wr, wi, leaf(p + "conv_r.bias"), leaf(p + "conv_i.bias"), tf32=tf and not tc16, tc=tc16))
            gp.add_packed_conv(self.enc_dgrad[i], TP.sym_dgrad_strided(wr, wi, STRIDE_E[i], tf32=tf))
            p = f"decoder.{i}." if i == Lr - 1 else f"decoder.{i}.0."
            wr, wi = leaf(p + "conv_tran_r.weight"), leaf(p + "conv_tran_i.weight")
            gp.add_packed_conv(self.dec[i], TP.sym_conv(wr, wi, leaf(p + "conv_tran_r.bias"), leaf(p + "conv_tran_i.bias"), transposed=True,
                                                        up=UPSAMPLE[i], tf32=tf and not tc16, tc=tc16))
            c0 = wr.shape[0] // 2
            for pk, sl in zip(self.dec_dgrad[i], (slice(0, c0), slice(c0, None))):
                gp.add_packed_conv(pk, TP.sym_dgrad(wr[sl], wi[sl], transposed=True, tf32=tf))
        s4 = off["fc.fc_r.weight"][1] + (1, 1)
        wr, wi = leaf("fc.fc_r.weight", s4), leaf("fc.fc_i.weight", s4)
        gp.add_packed_conv(self.fc, TP.sym_conv(wr, wi, leaf("fc.fc_r.bias"), leaf("fc.fc_i.bias"), tf32=tf))
        gp.add_packed_conv(self.fc_dgrad, TP.sym_dgrad(wr, wi, transposed=False, tf32=tf))
        names, sfx = ("real_lstm", "imag_lstm"), ("", "_reverse")
        for l in range(2):
            for j, n in enumerate(names):
                gp.add(TP.Sym.cat([leaf(f"lstm.{n}.weight_ih_l{l}{s}") for s in sfx], 0), torch.float32,
                       lambda t, l=l, j=j: self.lstm_wih[l].__setitem__(j, t), current=self.lstm_wih[l][j])
                gp.add(TP.Sym.cat([leaf(f"lstm.{n}.bias_ih_l{l}{s}") + leaf(f"lstm.{n}.bias_hh_l{l}{s}") for s in sfx], 0), torch.float32,
                       lambda t, l=l, j=j: self.lstm_b[l].__setitem__(j, t), current=self.lstm_b[l][j])
            gp.add(TP.Sym.stack([TP.Sym.stack([leaf(f"lstm.{n}.weight_hh_l{l}{s}") for s in sfx], 0) for n in names], 0), torch.float32,
                   lambda t, l=l: self.lstm_whh.__setitem__(l, t), current=self.lstm_whh[l])
        params = dict(self.model.named_parameters())
        def attention(ca_list, sa_list, prefix, count):
            for i in range(count):
                pc, ps = f"{prefix}.{2 * i}.", f"{prefix}.{2 * i + 1}."
                ca = ca_list[i]
                for key, name in (("w1_r", "fc.0.conv_r.weight"), ("w1_i", "fc.0.conv_i.weight"), ("w2_r", "fc.2.conv_r.weight"),
                                  ("w2_i", "fc.2.conv_i.weight")):
                    ca[key] = params[pc + name].data.view(ca[key].shape)               # views of the live parameters: nothing to re-pack
                gp.add(TP.Sym.cat([leaf(ps + "conv1.conv_r.weight").reshape(-1), leaf(ps + "conv1.conv_i.weight").reshape(-1)], 0), torch.float32,
                       lambda t, i=i: sa_list.__setitem__(i, t), current=sa_list[i])
        attention(self.skip_ca, self.skip_sa, "skip_attention", Lr)
        attention(self.dec_ca, self.dec_sa, "decoder_attention", Lr - 1)
        if str(dev).startswith("cpu"):        # CPU: tables only (tests check them against the host packing); there is no CPU gather kernel
            self.gather_tables = gp
            return
        self.gather = gp.finalize()
        self.gather.run()

    def optimizer_step(self, group=None):
        """Gradient exchange (NCCL all-reduce of the flat buckets when torch.distributed is initialised), global-norm clip
        (config.py:48-49) and Adam-amsgrad with L2 weight decay (c_network.py:229-235) on the flat buffers: two kernels."""
        if self.opt is None:
            self.init_optimizer()
        import torch.distributed as dist
        o, hp = self.opt, self.hp
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            for i in range(len(self.buckets.buckets)):
                self.buckets.launch(i, group)
            self.buckets.wait()
        flat = self.buckets.flat
        with torch.cuda.device(flat.device):
            o["step"] += 1
            lib = L.lib()
            L.check(lib.dcs_sumsq(L.ptr(flat), flat.numel(), L.ptr(o["sumsq"]), 0, L.ptr(o["ws"]), o["ws"].numel(), L.stream_ptr()), "dcs_sumsq")
            L.check(lib.dcs_adam_amsgrad(L.ptr(self.flat_param), L.ptr(flat), L.ptr(o["m"]), L.ptr(o["v"]), L.ptr(o["vmax"]), flat.numel(),
                                         float(hp.get("lr", 10e-5)), ADAM_BETAS[0], ADAM_BETAS[1], float(hp.get("optim_eps", 10e-7)),
                                         float(hp.get("optim_weight_decay", 10e-5)), o["step"], L.ptr(o["sumsq"]),
                                         float(hp.get("gradient_clip_val", 100.0)), 1.0 / world, L.stream_ptr()), "dcs_adam_amsgrad")
        self.steps_done += 1
        self._stale = True               # the packed operands are rebuilt (on the GPU) by the next forward
        return o["sumsq"]

    def step(self, noise_spec, noisy_spec, clean_spec):
        """forward + backward + optimizer step (c_network.py:243-261 training_step followed by Lightning's optimizer step)."""
        out = self.forward(noise_spec, noisy_spec, clean_spec)
        self.backward()
        self.optimizer_step()
        return out
