"""Standalone fp32 torch-CPU restatement of the DCS-Net hot path (TEST INFRASTRUCTURE ONLY).

This file is the parity oracle and the `cpu_baseline` arm; it is never the product path and must not be
imported from `dcs-net_b200/`.  It exists because `/root/reference` cannot travel to the GPU box.
PARITY UNPINNED: the reference has no tests/golden vectors for this path; this restatement is pinned in the
build container against the reference's own files run unmodified (`tests/test_oracle_vs_reference.py`) and
through the committed fixtures in `tests/golden/` (`oracle/make_golden.py`).

It is a *functional* restatement driven by a reference-format `state_dict`, and keeps the
reference's operation sequence (4 real convolutions per complex convolution, separate BN / activation / upsample /
concat passes, 4 LSTM passes) so that timing it on host cores is representative of the reference CPU path.

Reference lines followed:
  STFT front-end ............ data.py:112-134 with config.py:72-77
  C_NETWORK.forward ......... c_network.py:187-226 (layers built at 100-166)
  ComplexLSTM ............... c_network.py:33-47
  Channel / spatial attention c_network.py:62-84, network_functions.py:107-138
  bound_cRM ................. network_functions.py:77-88      complex_mat_mult: 90-96
  combine (dcs / dc) ........ network_functions.py:393-401 / 431-436
  mag_phase_2_wave .......... network_functions.py:140-150
  complexPyTorch 0.3 ........ third-party, restated per SURVEY Appendix A
"""
import math

import torch
import torch.nn.functional as F

# config.py:31-53, 72-106 — the values are the contract (the file itself is not copied).
HPARAMS = dict(no_of_layers=7, channels=[1, 16, 32, 64, 128, 256, 256, 256], lstm_layers=2, lstm_bidir=True,
               atan2_eps=10e-7, channel_attention_reduction_ratio=16, spatial_attention_kernel_size=7,
               dropout_conv=0.1, dropout_fc=0.2)
N_FFT, HOP, WIN = 512, 32, 512
KERNEL_E = [7, 7, 5, 5, 3, 3, 3]
STRIDE_E = [(2, 2), (2, 2), (2, 2), (2, 1), (2, 1), (2, 1), (2, 1)]
KERNEL_D = [3] * 7
UPSAMPLE = [(2, 1), (2, 1), (2, 1), (2, 1), (2, 2), (2, 2), (2, 2)]
BN_EPS = 1e-5
LRELU_SLOPE = 0.01
SAMPLE_RATE = 16000


def _cplx(re, im):
    return re.type(torch.complex64) + 1j * im.type(torch.complex64)


# ---------------------------------------------------------------- complexPyTorch 0.3 (Appendix A)
def apply_complex(fr, fi, x):
    return _cplx(fr(x.real) - fi(x.imag), fr(x.imag) + fi(x.real))


def cconv2d(x, sd, p, stride=1, padding=0):
    wr, wi = sd[p + "conv_r.weight"], sd[p + "conv_i.weight"]
    br, bi = sd.get(p + "conv_r.bias"), sd.get(p + "conv_i.bias")
    return apply_complex(lambda t: F.conv2d(t, wr, br, stride, padding),
                         lambda t: F.conv2d(t, wi, bi, stride, padding), x)


def cconvT2d(x, sd, p, stride=1, padding=0):
    wr, wi = sd[p + "conv_tran_r.weight"], sd[p + "conv_tran_i.weight"]
    br, bi = sd.get(p + "conv_tran_r.bias"), sd.get(p + "conv_tran_i.bias")
    return apply_complex(lambda t: F.conv_transpose2d(t, wr, br, stride, padding),
                         lambda t: F.conv_transpose2d(t, wi, bi, stride, padding), x)


def clinear(x, sd, p):
    return apply_complex(lambda t: F.linear(t, sd[p + "fc_r.weight"], sd[p + "fc_r.bias"]),
                         lambda t: F.linear(t, sd[p + "fc_i.weight"], sd[p + "fc_i.bias"]), x)


def cbn_eval(x, sd, p, eps=BN_EPS):
    """ComplexBatchNorm2d in eval mode: centre, 2x2 whitening from running_covar(+eps), 2x2 affine."""
    b = lambda v: v[None, :, None, None]
    x = x - b(sd[p + "running_mean"])
    cov = sd[p + "running_covar"]
    Crr, Cii, Cri = cov[:, 0] + eps, cov[:, 1] + eps, cov[:, 2]
    s = torch.sqrt(Crr * Cii - Cri.pow(2))
    t = torch.sqrt(Cii + Crr + 2 * s)
    ist = 1.0 / (s * t)
    Rrr, Rii, Rri = (Cii + s) * ist, (Crr + s) * ist, -Cri * ist
    x = _cplx(b(Rrr) * x.real + b(Rri) * x.imag, b(Rii) * x.imag + b(Rri) * x.real)
    w, c = sd[p + "weight"], sd[p + "bias"]
    return _cplx(b(w[:, 0]) * x.real + b(w[:, 2]) * x.imag + b(c[:, 0]),
                 b(w[:, 2]) * x.real + b(w[:, 1]) * x.imag + b(c[:, 1]))


def crelu(x):
    return _cplx(F.relu(x.real), F.relu(x.imag))


def clrelu(x):  # network_functions.py:104-105, default negative_slope 0.01
    return torch.complex(F.leaky_relu(x.real), F.leaky_relu(x.imag))


def csigmoid(x):  # network_functions.py:111-112
    return _cplx(torch.sigmoid(x.real), torch.sigmoid(x.imag))


def cupsample_nearest(x, scale):
    return _cplx(F.interpolate(x.real, scale_factor=scale, mode="nearest"),
                 F.interpolate(x.imag, scale_factor=scale, mode="nearest"))


# ---------------------------------------------------------------- c_network.py blocks
def channel_attention(x, sd, p):
    """c_network.py:62-69.  The 'max' pool is an average pool (network_functions.py:135-138) => 2*fc(avg)."""
    pool = lambda t: _cplx(F.adaptive_avg_pool2d(t.real, 1), F.adaptive_avg_pool2d(t.imag, 1))
    fc = lambda t: cconv2d(crelu(cconv2d(t, sd, p + "fc.0.")), sd, p + "fc.2.")
    return csigmoid(fc(pool(x)) + fc(pool(x)))


def spatial_attention(x, sd, p, k=7):
    """c_network.py:77-84."""
    avg = torch.mean(x, dim=1, keepdim=True)
    mx = torch.complex(torch.max(x.real, dim=1, keepdim=True)[0], torch.max(x.imag, dim=1, keepdim=True)[0])
    return csigmoid(cconv2d(torch.cat([avg, mx], dim=1), sd, p + "conv1.", padding=k // 2))


def _lstm_stack(x, sd, p, layers=2, hidden=64):
    """torch.nn.LSTM(batch_first, bidirectional) forward, zero initial state, gate order i,f,g,o."""
    B, S, _ = x.shape
    inp = x
    for l in range(layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            wih, whh = sd[f"{p}weight_ih_l{l}{suffix}"], sd[f"{p}weight_hh_l{l}{suffix}"]
            bias = sd[f"{p}bias_ih_l{l}{suffix}"] + sd[f"{p}bias_hh_l{l}{suffix}"]
            pre = inp @ wih.t() + bias
            h = x.new_zeros(B, hidden)
            c = x.new_zeros(B, hidden)
            hs = [None] * S
            order = range(S - 1, -1, -1) if rev else range(S)
            for t in order:
                g = pre[:, t] + h @ whh.t()
                i, f, gg, o = g.split(hidden, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                hs[t] = h
            outs.append(torch.stack(hs, dim=1))
        inp = torch.cat(outs, dim=2)
    return inp


def _lstm_module(sd, p, in_dim, hidden, layers):
    """Same arithmetic through torch.nn.LSTM (used for timing fidelity: the reference calls nn.LSTM)."""
    m = torch.nn.LSTM(in_dim, hidden, num_layers=layers, bidirectional=True, batch_first=True)
    m.load_state_dict({k[len(p):]: v for k, v in sd.items() if k.startswith(p)})
    return m.eval()


def complex_lstm(x, sd, p, layers=2, hidden=64, explicit=False, _cache={}):
    """c_network.py:33-47: four real LSTM passes, out = (R(re) - I(im)) + j (R(im) + I(re))."""
    if explicit:
        R = lambda t: _lstm_stack(t, sd, p + "real_lstm.", layers, hidden)
        I = lambda t: _lstm_stack(t, sd, p + "imag_lstm.", layers, hidden)
    else:
        key = (id(sd), p)
        if key not in _cache:
            _cache.clear()
            _cache[key] = (_lstm_module(sd, p + "real_lstm.", x.shape[-1], hidden, layers),
                           _lstm_module(sd, p + "imag_lstm.", x.shape[-1], hidden, layers))
        rl, il = _cache[key]
        R, I = (lambda t: rl(t)[0]), (lambda t: il(t)[0])
    re, im = x.real, x.imag
    r2r, r2i, i2r, i2i = R(re), I(re), R(im), I(im)
    return torch.complex(r2r - i2i, i2r + r2i)


def bound_crm(m, eps=HPARAMS["atan2_eps"]):
    """network_functions.py:77-88 — tanh-bounded magnitude, phase re-derived twice with +eps on the real part."""
    t = torch.tanh(torch.abs(m))
    th1 = torch.atan2(m.imag, m.real + eps)
    r1, i1 = t * torch.cos(th1), t * torch.sin(th1)
    th2 = torch.atan2(i1, r1 + eps)
    return torch.complex(t * torch.cos(th2), t * torch.sin(th2))


def c_network_forward(sd, x, hp=HPARAMS, taps=None, explicit_lstm=False, bn=None, drop=None):
    """C_NETWORK.forward in eval mode (c_network.py:187-226).  x: (B,F,T) complex64 -> bounded mask, squeezed.

    `taps`, if a dict, receives named intermediate activations (NCHW complex64) for per-layer parity tests.
    `bn(x, sd, prefix)` / `drop(x, kind)` replace the eval-mode batch norm / the identity dropout for the train-mode
    restatement in oracle/train_oracle.py (dropout sits at c_network.py:195, 203, 221).
    """
    L = hp["no_of_layers"]
    tap = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
    cbn_eval = bn or globals()["cbn_eval"]
    drop = drop or (lambda t, kind: t)
    e = cbn_eval(x.view(x.shape[0], -1, x.shape[1], x.shape[2]), sd, "initial_batchnorm.")
    tap("bn0", e)
    enc = [e]
    for i in range(L):
        k = KERNEL_E[i]
        e = cconv2d(enc[i], sd, f"encoder.{i}.0.", STRIDE_E[i], k // 2)
        e = crelu(cbn_eval(e, sd, f"encoder.{i}.1."))
        tap(f"enc{i}", e)
        e = drop(e, "conv")
        enc.append(e)  # dropout is the identity in eval mode (c_network.py:195-196)
    shp = enc[-1].shape
    seq = torch.flatten(e, 2, 3).permute(0, 2, 1)
    lo = complex_lstm(seq, sd, "lstm.", hp["lstm_layers"], hp["channels"][4] // 2, explicit=explicit_lstm)
    tap("lstm", lo)
    fo = clinear(lo, sd, "fc.")
    tap("fc", fo)
    fo = drop(fo, "fc")
    d = fo.permute(0, 2, 1).reshape(shp)
    for i in range(L):
        skip = enc[L - i]
        ca = channel_attention(skip, sd, f"skip_attention.{2 * i}.") * skip
        sa = spatial_attention(ca, sd, f"skip_attention.{2 * i + 1}.", hp["spatial_attention_kernel_size"]) * ca
        tap(f"skip{i}", sa)
        d = torch.cat((d, sa), dim=1)
        d = cupsample_nearest(d, UPSAMPLE[i])
        if i == L - 1:
            d = cconvT2d(d, sd, f"decoder.{i}.", 1, KERNEL_D[i] // 2)
        else:
            d = cconvT2d(d, sd, f"decoder.{i}.0.", 1, KERNEL_D[i] // 2)
            d = clrelu(cbn_eval(d, sd, f"decoder.{i}.1."))
            tap(f"dec{i}_act", d)
            d = d * channel_attention(d, sd, f"decoder_attention.{2 * i}.")
            d = d * spatial_attention(d, sd, f"decoder_attention.{2 * i + 1}.", hp["spatial_attention_kernel_size"])
        tap(f"dec{i}", d)
        d = drop(d, "conv")
    out = torch.squeeze(d)  # also drops the batch dim at B=1 (Appendix D6)
    return bound_crm(out, hp["atan2_eps"])


# ---------------------------------------------------------------- front / back end
def stft(audio):
    """data.py:112-134: centre reflect-pad, hann(512) periodic, hop 32, normalized, bins 1..256 kept."""
    spec = torch.stft(audio, n_fft=N_FFT, hop_length=HOP, win_length=WIN, window=torch.hann_window(WIN),
                      return_complex=True, normalized=True)
    return spec[..., 1:N_FFT // 2 + 1, :]


def spec_to_wave(s, eps=HPARAMS["atan2_eps"]):
    """abs/atan2(+eps) polar split (network_functions.py:398-401) then mag_phase_2_wave (140-150): zero row
    appended at the END of the frequency axis, torch.istft(normalized)."""
    mag, ph = torch.abs(s), torch.atan2(s.imag, s.real + eps)
    comp = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))
    comp = F.pad(comp, (0, 0, 0, 1))
    return torch.istft(comp, n_fft=N_FFT, hop_length=HOP, win_length=WIN, window=torch.hann_window(WIN),
                       normalized=True)


def enhance_spec(sd, noisy_spec, variant="dcs", hp=HPARAMS, taps=None):
    """network_functions.py:393-397 (dcs) / 431-434 (dc): second bound_cRM, complex product, subtraction."""
    with torch.no_grad():
        net_out = c_network_forward(sd, noisy_spec, hp, taps)
        mask = bound_crm(net_out, hp["atan2_eps"])
        prod = torch.complex(noisy_spec.real * mask.real - noisy_spec.imag * mask.imag,
                             noisy_spec.real * mask.imag + noisy_spec.imag * mask.real)
        if variant in ("dcs", "drs"):
            noise_spec, clean_spec = prod, noisy_spec - prod
        else:
            noise_spec, clean_spec = None, prod
    return dict(net_out=net_out, mask=mask, noise_spec=noise_spec, clean_spec=clean_spec)


def enhance_audio(sd, noisy_audio, variant="dcs", hp=HPARAMS):
    """Full hot path on CPU: STFT -> net -> bound -> combine -> iSTFT.  (B,L) fp32 -> dict incl. clean_audio."""
    with torch.no_grad():
        spec = stft(noisy_audio)
        r = enhance_spec(sd, spec, variant, hp)
        r["noisy_spec"] = spec
        r["clean_audio"] = spec_to_wave(r["clean_spec"], hp["atan2_eps"])
        if r["noise_spec"] is not None:
            r["noise_audio"] = spec_to_wave(r["noise_spec"], hp["atan2_eps"])
    return r


def si_snr(clean, estimate, eps=1e-8):
    """network_functions.py:30-42 (the parity metric for |dSI-SDR| <= 0.01 dB)."""
    dot = torch.sum(estimate * clean, -1, keepdim=True)
    norm = torch.sum(clean * clean, -1, keepdim=True)
    s_t = dot * clean / (norm + eps)
    e_n = estimate - s_t
    snr = 10 * torch.log10(torch.sum(s_t * s_t, -1, keepdim=True) / (torch.sum(e_n * e_n, -1, keepdim=True) + eps) + eps)
    return torch.mean(snr)


# ---------------------------------------------------------------- synthetic inputs
def synthetic_audio(batch, length, seed=1234):
    g = torch.Generator().manual_seed(seed)
    clean = 0.1 * torch.randn(batch, length, generator=g)
    noise = 0.05 * torch.randn(batch, length, generator=g)
    t = torch.arange(length, dtype=torch.float32) / SAMPLE_RATE
    for f0, a in ((220.0, 0.08), (1330.0, 0.05), (3100.0, 0.03)):  # spectral structure
        clean = clean + a * torch.sin(2 * math.pi * f0 * t)[None, :]
    return clean, noise, clean + noise


def audio_seconds(n_frames):
    return HOP * (n_frames - 1) / SAMPLE_RATE


def frames_to_samples(n_frames):
    return HOP * (n_frames - 1)
