"""Drop-in twin of the model classes of the reference's r_network.py (dr / drs): RealChannelAttention,
RealSpatialAttention and R_NETWORK(config, hparams, seed) — same constructor signatures, attribute names and registration
order, hence the same 158 state_dict keys / shapes and, for a given seed, bit-identical random-init weights (pinned by
tests/test_rnet_oracle.py against tests/golden/rnet_*.pt, which were produced by the reference's own r_network.py).

`R_NETWORK.forward` runs in eval mode in two precisions, selected by the (non-reference) attribute `compute_mode`:
  'fp32' (default, <= 1e-5): a sequence of sm_100a CUDA-core kernels — the BatchNorm'd magnitude (dcs_cbn_apply), every
      Conv2d / ConvTranspose2d + BatchNorm2d + activation on the fp32 conv kernel through real packing (packing.PackedRNet:
      real channels 2c, 2c+1 <-> (re, im) of a channel pair; cat + nearest up-sampling folded into the decoder GEMMs; the
      final sigmoid in the last epilogue), the real CBAM (dcs_real_attention_fwd) and the real LSTM (dcs_rlstm_fwd);
  'fp16' / 'bf16' (<= 2e-3): the tensor-core plan of rengine.RealForwardPlan (tcgen05 convs, fp16 mma.sync LSTM, streaming
      CBAM, fused sigmoid tail).  The audio -> audio hot path with CUDA-graph replay is dcsnet_b200.RealEnhancer.
No ATen / CPU fallback: CPU tensors and train mode raise.  The stand-alone attention modules are parameter containers (their
forward raises).
"""
import torch

from . import ops, packing
from .c_network import _Base, _StepMixin, _seed_everything


class RealChannelAttention(torch.nn.Module):
    """r_network.py:8-25 (only the max-pool branch reaches the output)."""

    def __init__(self, no_channels, reduction_ratio):
        super().__init__()
        self.avg_pool = torch.nn.AdaptiveAvgPool2d(1)
        self.max_pool = torch.nn.AdaptiveMaxPool2d(1, return_indices=False)
        self.fc = torch.nn.Sequential(torch.nn.Conv2d(no_channels, max(no_channels // reduction_ratio, 1), 1, bias=False),
                                      torch.nn.ReLU(),
                                      torch.nn.Conv2d(max(no_channels // reduction_ratio, 1), no_channels, 1, bias=False))
        self.sigmoid = torch.nn.Sigmoid()

    def forward(self, x):
        raise NotImplementedError("dcsnet_b200: the real (dr / drs) path has no sm_100a kernels yet (SURVEY 8f rank 1); no CPU fallback")


class RealSpatialAttention(torch.nn.Module):
    """r_network.py:28-40."""

    def __init__(self, kernel_size):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = torch.nn.Sigmoid()

    def forward(self, x):
        raise NotImplementedError("dcsnet_b200: the real (dr / drs) path has no sm_100a kernels yet (SURVEY 8f rank 1); no CPU fallback")


class R_NETWORK(_StepMixin, _Base):
    """r_network.py:43-173."""
    _step_dtype = "real"

    def __init__(self, config, hparams, seed):
        super().__init__()
        _seed_everything(seed)
        self.config = config
        self.hparams.update(hparams)
        self.save_hyperparameters(self.hparams)
        hp = self.hparams
        self.encoder = torch.nn.ModuleList()
        self.decoder = torch.nn.ModuleList()
        self.decoder_attention = torch.nn.ModuleList()
        self.skip_attention = torch.nn.ModuleList()
        n_layers, ch = hp['no_of_layers'], hp['channels']
        self.initial_batchnorm = torch.nn.BatchNorm2d(ch[0])
        for i in range(n_layers):
            self.encoder.append(torch.nn.Sequential(
                torch.nn.Conv2d(in_channels=1 if i == 0 else ch[i], out_channels=ch[i + 1], kernel_size=config.kernel_sizeE[i],
                                stride=config.strideE[i], padding=config.paddingE[i]),
                torch.nn.BatchNorm2d(ch[i + 1]),
                config.RactivationE()))
        self.lstm = torch.nn.LSTM(input_size=ch[5], hidden_size=ch[4], num_layers=hp['lstm_layers'],
                                  bidirectional=hp['lstm_bidir'], batch_first=True)
        self.fc = torch.nn.Linear(ch[5], ch[5])
        self.dropout_conv = torch.nn.Dropout(hp['dropout_conv'])
        self.dropout_fc = torch.nn.Dropout(hp['dropout_fc'])
        for i in range(n_layers):
            in_channels = ch[n_layers - i]
            out_channels = max(ch[n_layers - 1 - i], 1)
            convt = torch.nn.ConvTranspose2d(in_channels + in_channels, out_channels, kernel_size=config.kernel_sizeD[i],
                                             stride=config.strideD, padding=config.paddingD[i])
            if i == n_layers - 1:
                self.decoder.append(convt)
            else:
                self.decoder.append(torch.nn.Sequential(convt, torch.nn.BatchNorm2d(ch[n_layers - 1 - i]), config.RactivationD()))
            self.skip_attention.append(RealChannelAttention(in_channels, hp['channel_attention_reduction_ratio']))
            self.skip_attention.append(RealSpatialAttention(hp['spatial_attention_kernel_size']))
            self.decoder_attention.append(RealChannelAttention(out_channels, hp['channel_attention_reduction_ratio']))
            self.decoder_attention.append(RealSpatialAttention(hp['spatial_attention_kernel_size']))
        self.weights_init()
        self.compute_mode = "fp32"
        self._plans = {}

    def _tc_plan(self, device, B, Fb, T):
        from .rengine import PackedRealNet, RealForwardPlan
        pkey = (str(device), self.compute_mode) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, "_tc_pk_key", None) != pkey:
            self._tc_pk, self._tc_pk_key = PackedRealNet(self.state_dict(), device, self.compute_mode, self.hparams['no_of_layers']), pkey
            self._plans = {}
        key = (B, Fb, T)
        if key not in self._plans:
            self._plans[key] = RealForwardPlan(self._tc_pk, B, T, n_bins=Fb, variant="drs")
        return self._plans[key]

    def weights_init(self):
        init = self.hparams['initialisation_distribution']
        for m in self.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
                init(m.weight)

    def _packed(self, device):
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, "_pk_key", None) != key:
            self._pk, self._pk_key = packing.PackedRNet(self.state_dict(), device=device, no_of_layers=self.hparams['no_of_layers']), key
        return self._pk

    def forward(self, x):
        """r_network.py:125-173: x (B, F, T) fp32 magnitude -> sigmoid mask, squeezed."""
        if self.training:
            raise NotImplementedError("dcsnet_b200.R_NETWORK: only the eval-mode forward is built (training step: SURVEY 8f rank 2)")
        if not x.is_cuda:
            raise RuntimeError("dcsnet_b200.R_NETWORK.forward needs CUDA tensors (sm_100a); there is no CPU fallback")
        B, Fb, T = x.shape
        if self.compute_mode != "fp32":      # tensor-core plan; the mask does not depend on the phase, so Y = |Y| + 0j serves
            plan = self._tc_plan(x.device, B, Fb, T)
            with torch.cuda.device(x.device):
                plan.Y.copy_(torch.complex(x.float(), torch.zeros_like(x, dtype=torch.float32)))
                plan._enqueue_from_spec()
            return torch.squeeze(plan.mask.clone())
        pk, Lr = self._packed(x.device), self.hparams['no_of_layers']
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=x.device)   # noqa: E731
        t = torch.zeros(B, Fb, T, 1, 2, dtype=torch.float32, device=x.device)
        t[..., 0, 0].copy_(x)                                     # magnitude in the pair's first slot, 0 in the padding slot
        enc = [ops.cbn_apply(t, pk.bn0)]                          # initial_batchnorm
        H, W = Fb, T
        for i in range(Lr):
            H, W = ops.conv_out_hw(pk.enc[i], H, W)
            enc.append(ops.cconv(pk.enc[i], enc[i], None, new(B, H, W, pk.enc[i].cout, 2)))
        lat = ops.rlstm(enc[-1].view(B, H * W, -1), pk.lstm_t)    # sequence index = h * W + w (flatten(2, 3).permute(0, 2, 1))
        d = ops.cconv(pk.fc, lat.view(B, 1, H * W, -1, 2), None, new(B, 1, H * W, pk.fc.cout, 2)).view(B, H, W, pk.fc.cout, 2)
        for i in range(Lr):
            skip = ops.real_attention(enc[Lr - i], *pk.skip_att[i])
            H, W = H * pk.dec[i].up[0], W * pk.dec[i].up[1]
            d = ops.cconv(pk.dec[i], d, skip, new(B, H, W, pk.dec[i].cout, 2))
            if i != Lr - 1:
                d = ops.real_attention(d, *pk.dec_att[i])
        return torch.squeeze(d[..., 0, 0])                        # sigmoid applied by the last conv's epilogue


def enhance_batch_real(network, noisy_spec, variant="drs", atan2_eps=10e-7):
    """The inference lines of the reference's real-path step functions (network_functions.py:286-305 drs, 338-342 dr) on the
    GPU: |Y|, noisy phase, mask = network(|Y|), magnitude combine, mag_phase_2_wave with the NOISY phase (iSTFT kernel)."""
    assert variant in ("dr", "drs")
    mag, phase = ops.mag_phase(noisy_spec, atan2_eps)
    mask = network(mag)
    if mask.dim() == 2:
        mask = mask[None]
    clean_mag, noise_mag = ops.real_mask_combine(mag, mask, subtract=variant == "drs")
    res = dict(predict_noise_mask=mask, predict_clean_mag=clean_mag, predict_noise_mag=noise_mag, noisy_mag=mag, noisy_phase=phase,
               predict_clean_audio=ops.istft_mag_phase(clean_mag, phase))
    if noise_mag is not None:                                     # drs: network_functions.py:304 / 390
        res["predict_noise_audio"] = ops.istft_mag_phase(noise_mag, phase)
    return res
