"""CPU restatement of the reference's TRAINING step, complex (dcs / dc) and real (drs / dr) networks (TEST INFRASTRUCTURE ONLY;
SURVEY 8f rank 2).

The product side is dcsnet_b200.train_engine.TrainStep; this is its oracle, pinned by tests/golden/train_step.pt (the reference's own
`train_batch_2_loss` + `backward()`, oracle/make_golden_train.py).  The forward is oracle/dcsnet_oracle.c_network_forward
with train-mode ComplexBatchNorm2d (batch statistics, running-stat update; complexPyTorch 0.3) and an
optional dropout hook; gradients come from torch autograd over this restatement.

Reference lines followed:
  train_batch_2_loss ........ network_functions.py:210-280 (dcs: 236-258, dc: 271-280)
  calc_loss ................. network_functions.py:168-208 (noise_loss_type 6, speech_loss_type 0 = config.py defaults)
  ComplexBatchNorm2d.train .. complexPyTorch/complexLayers.py (momentum 0.1, unbiased running covariance)
  dropout positions ......... c_network.py:195, 203, 221 (on view_as_real, i.e. real and imaginary parts drop independently)
"""
import math

import torch

from . import dcsnet_oracle as O

BN_MOMENTUM = 0.1


def cbn_train(new_stats):
    """Returns bn(x, sd, prefix): batch-statistic complex whitening + affine; the updated running statistics go to `new_stats`."""
    def bn(x, sd, p, eps=O.BN_EPS):
        b = lambda v: v[None, :, None, None]  # noqa: E731
        mean = torch.complex(x.real.mean([0, 2, 3]), x.imag.mean([0, 2, 3]))
        x = x - b(mean)
        n = x.numel() / x.size(1)
        Crr = x.real.pow(2).sum(dim=[0, 2, 3]) / n + eps
        Cii = x.imag.pow(2).sum(dim=[0, 2, 3]) / n + eps
        Cri = (x.real * x.imag).mean(dim=[0, 2, 3])
        with torch.no_grad():
            m = BN_MOMENTUM
            new_stats[p + "running_mean"] = m * mean + (1 - m) * sd[p + "running_mean"]
            cov = sd[p + "running_covar"]
            new_stats[p + "running_covar"] = torch.stack([m * Crr * n / (n - 1) + (1 - m) * cov[:, 0],
                                                          m * Cii * n / (n - 1) + (1 - m) * cov[:, 1],
                                                          m * Cri * n / (n - 1) + (1 - m) * cov[:, 2]], dim=1)
        s = torch.sqrt(Crr * Cii - Cri.pow(2))
        t = torch.sqrt(Cii + Crr + 2 * s)
        ist = 1.0 / (s * t)
        Rrr, Rii, Rri = (Cii + s) * ist, (Crr + s) * ist, -Cri * ist
        re, im = b(Rrr) * x.real + b(Rri) * x.imag, b(Rii) * x.imag + b(Rri) * x.real
        w, c = sd[p + "weight"], sd[p + "bias"]
        return torch.complex(b(w[:, 0]) * re + b(w[:, 2]) * im + b(c[:, 0]), b(w[:, 2]) * re + b(w[:, 1]) * im + b(c[:, 1]))
    return bn


def dropout_from_masks(masks):
    """drop(x, kind) applying caller-supplied keep masks in call order (scaled 1 / (1 - p) by the caller); None = identity."""
    it = iter(masks)

    def drop(x, kind):
        m = next(it)
        return x if m is None else torch.view_as_complex(torch.view_as_real(x) * m)
    return drop


def dropout_torch_stream(p_conv, p_fc):
    """drop(x, kind) drawing from torch's global CPU generator exactly as the reference does (torch.nn.Dropout on
    view_as_real, c_network.py:195-196 / 203-204 / 221-222): after the same torch.manual_seed the masks are identical."""
    import torch.nn.functional as F

    def drop(x, kind):
        return torch.view_as_complex(F.dropout(torch.view_as_real(x), p_fc if kind == "fc" else p_conv, True))
    return drop


def _mul(a, b):
    return torch.complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def train_step(sd, noise_spec, noisy_spec, clean_spec, param_names, variant="dcs", hp=O.HPARAMS, speech_alpha=0.7, drop=None):
    """One training step's forward + backward.  `sd`: reference-format state_dict (not modified); `param_names`: the keys
    that are nn.Parameters (the rest are buffers).  Returns dict(noise_loss, speech_loss, train_loss, grads{name: tensor},
    running_stats{name: tensor})."""
    eps = hp["atan2_eps"]
    live = {k: (v.detach().clone().requires_grad_(True) if k in param_names else v) for k, v in sd.items()}
    stats = {}
    mask_out = O.c_network_forward(live, noisy_spec, hp, explicit_lstm=True, bn=cbn_train(stats), drop=drop)
    if mask_out.dim() == 2:
        mask_out = mask_out[None]
    mask = O.bound_crm(mask_out, eps)                                   # second bound, network_functions.py:240 / 273
    prod = _mul(noisy_spec, mask)
    wave = lambda s: O.spec_to_wave(s, eps)                             # noqa: E731
    clean_audio = wave(clean_spec)
    if variant == "dcs":
        noise_loss = 1 - speech_alpha * (-O.si_snr(wave(noise_spec), wave(prod)))        # line 195-196 (precedence as written)
        speech_loss = speech_alpha * (-O.si_snr(clean_audio, wave(noisy_spec - prod)))
        total = noise_loss + speech_loss
    else:
        noise_loss = None
        speech_loss = speech_alpha * (-O.si_snr(clean_audio, wave(prod)))
        total = speech_loss
    total.backward()
    return dict(noise_loss=None if noise_loss is None else float(noise_loss.detach()), speech_loss=float(speech_loss.detach()),
                train_loss=float(total.detach()), grads={k: live[k].grad for k in param_names if live[k].grad is not None},
                running_stats=stats)


def train_step_real(sd, noise_spec, noisy_spec, clean_spec, param_names, variant="drs", speech_alpha=0.7, atan2_eps=10e-7):
    """The real path's training step (network_functions.py:223-233 drs, 260-267 dr; r_network.py:125-173 in train mode)."""
    from . import rnet_oracle as RO
    live = {k: (v.detach().clone().requires_grad_(True) if k in param_names else v) for k, v in sd.items()}
    stats = {}
    mag = lambda s: torch.abs(s)                                          # noqa: E731
    wave = lambda s: O.spec_to_wave(s, atan2_eps)                         # noqa: E731
    noisy_mag, noisy_phase = mag(noisy_spec), torch.atan2(noisy_spec.imag, noisy_spec.real + atan2_eps)
    mask = RO.r_network_forward(live, noisy_mag, train_stats=stats)
    if mask.dim() == 2:
        mask = mask[None]
    clean_audio = wave(clean_spec)
    if variant == "drs":
        noise_mag = noisy_mag * mask
        noise_loss = 1 - speech_alpha * (-O.si_snr(wave(noise_spec), RO.mag_phase_2_wave(noise_mag, noisy_phase)))
        speech_loss = speech_alpha * (-O.si_snr(clean_audio, RO.mag_phase_2_wave(noisy_mag - noise_mag, noisy_phase)))
        total = noise_loss + speech_loss
    else:
        noise_loss = None
        speech_loss = speech_alpha * (-O.si_snr(clean_audio, RO.mag_phase_2_wave(noisy_mag * mask, noisy_phase)))
        total = speech_loss
    total.backward()
    return dict(noise_loss=None if noise_loss is None else float(noise_loss.detach()), speech_loss=float(speech_loss.detach()),
                train_loss=float(total.detach()), grads={k: live[k].grad for k in param_names if live[k].grad is not None},
                running_stats=stats)


def istft_adjoint(g, n_frames):
    """Closed form of d<g, mag_phase_2_wave(S)>/dS for the reference's iSTFT (network_functions.py:140-150: zero row appended
    at the END of the frequency axis, n_fft 512, hop 32, hann, normalized, centre-trimmed) — the backward kernel's contract.

    g: (B, 32 (T-1)) waveform gradient -> (B, 256, T) complex64 gradient in torch's convention (dL/dRe + j dL/dIm):
    zero-pad g by 256 on both sides, divide by the overlap-add envelope sum_t w^2, then a NON-centred, normalized STFT with the
    same window; rfft bins 0..255 are kept (the forward put spectrogram row k on rfft bin k), scaled by 2 for bins 1..255
    (Hermitian halves of the C2R transform) and by 1 with the imaginary part dropped for bin 0 (C2R ignores it).
    So it is the forward STFT kernel with: no reflect padding, a 1/envelope pre-scale, bin offset 0 instead of 1, and the
    per-bin factor."""
    n, hop = O.N_FFT, O.HOP
    win = torch.hann_window(n)
    full = hop * (n_frames - 1) + n
    env = torch.zeros(full)
    for t in range(n_frames):
        env[t * hop:t * hop + n] += win ** 2
    G = torch.nn.functional.pad(g, (n // 2, n // 2)) / env.clamp_min(1e-20)
    X = torch.stft(G, n_fft=n, hop_length=hop, win_length=n, window=win, center=False, normalized=True, return_complex=True)[:, :256]
    scale = torch.full((256,), 2.0)
    scale[0] = 1.0
    X = X * scale[None, :, None]
    imag = X.imag.clone()
    imag[:, 0] = 0.0
    return torch.complex(X.real, imag)


def cbn_train_backward(x, dy, weight, eps=O.BN_EPS):
    """Closed-form backward of train-mode ComplexBatchNorm2d (the contract of the BN backward kernel): x, dy complex (B,C,H,W)
    with dy = dL/dRe y + j dL/dIm y; weight (C,3).  Returns (dx complex, dweight (C,3), dbias (C,2)).

    Two reduction passes per channel, like the forward: pass 1 accumulates the eight sums below (dW, db, and <dz, xc> for the
    whitening matrix), a per-channel 3x3 Jacobian turns dR into dC, pass 2 forms dx = R^T dz + (2/n)(dC applied to xc) and
    removes its mean."""
    b = lambda v: v[None, :, None, None]  # noqa: E731
    red = lambda v: v.sum(dim=[0, 2, 3])  # noqa: E731
    xr, xi = x.real - b(x.real.mean([0, 2, 3])), x.imag - b(x.imag.mean([0, 2, 3]))
    n = x.numel() / x.size(1)
    A, Bc, Cc = red(xr * xr) / n + eps, red(xi * xi) / n + eps, red(xr * xi) / n
    s = torch.sqrt(A * Bc - Cc * Cc)
    t = torch.sqrt(A + Bc + 2 * s)
    u = 1.0 / (s * t)
    Rrr, Rii, Rri = (Bc + s) * u, (A + s) * u, -Cc * u
    zr, zi = b(Rrr) * xr + b(Rri) * xi, b(Rii) * xi + b(Rri) * xr
    w0, w1, w2 = weight[:, 0], weight[:, 1], weight[:, 2]
    gr, gi = dy.real, dy.imag
    dweight = torch.stack([red(gr * zr), red(gi * zi), red(gr * zi + gi * zr)], dim=1)
    dbias = torch.stack([red(gr), red(gi)], dim=1)
    dzr, dzi = b(w0) * gr + b(w2) * gi, b(w2) * gr + b(w1) * gi
    dRrr, dRii, dRri = red(dzr * xr), red(dzi * xi), red(dzr * xi + dzi * xr)
    # Jacobian of (Rrr, Rii, Rri) w.r.t. (A, B, C)
    s_a, s_b, s_c = Bc / (2 * s), A / (2 * s), -Cc / s
    t_a, t_b, t_c = (1 + 2 * s_a) / (2 * t), (1 + 2 * s_b) / (2 * t), s_c / t
    u_a, u_b, u_c = -u * (s_a / s + t_a / t), -u * (s_b / s + t_b / t), -u * (s_c / s + t_c / t)
    dA = dRrr * (s_a * u + (Bc + s) * u_a) + dRii * ((1 + s_a) * u + (A + s) * u_a) + dRri * (-Cc * u_a)
    dB = dRrr * ((1 + s_b) * u + (Bc + s) * u_b) + dRii * (s_b * u + (A + s) * u_b) + dRri * (-Cc * u_b)
    dC = dRrr * (s_c * u + (Bc + s) * u_c) + dRii * (s_c * u + (A + s) * u_c) + dRri * (-u - Cc * u_c)
    dxr = b(Rrr) * dzr + b(Rri) * dzi + (2 * b(dA) * xr + b(dC) * xi) / n
    dxi = b(Rii) * dzi + b(Rri) * dzr + (2 * b(dB) * xi + b(dC) * xr) / n
    dxr, dxi = dxr - b(dxr.mean([0, 2, 3])), dxi - b(dxi.mean([0, 2, 3]))
    return torch.complex(dxr, dxi), dweight, dbias


def bound_crm_backward(m, dout, eps=O.HPARAMS["atan2_eps"]):
    """Closed-form backward of bound_cRM (network_functions.py:77-88; applied twice on the training path, c_network.py:225 and
    network_functions.py:240): m complex, dout = dL/dRe out + j dL/dIm out -> dL/dm in the same convention.  Element-wise."""
    mr, mi, gr, gi = m.real, m.imag, dout.real, dout.imag
    rho = torch.sqrt(mr * mr + mi * mi)
    t = torch.tanh(rho)
    a = mr + eps
    th1 = torch.atan2(mi, a)
    c1, s1 = torch.cos(th1), torch.sin(th1)
    r1, i1 = t * c1, t * s1
    a2 = r1 + eps
    th2 = torch.atan2(i1, a2)
    c2, s2 = torch.cos(th2), torch.sin(th2)
    dt = gr * c2 + gi * s2
    dth2 = t * (-gr * s2 + gi * c2)
    q2 = a2 * a2 + i1 * i1
    dr1, di1 = dth2 * (-i1 / q2), dth2 * (a2 / q2)
    dt = dt + dr1 * c1 + di1 * s1
    dth1 = t * (-dr1 * s1 + di1 * c1)
    q1 = a * a + mi * mi
    sech2 = 1 - t * t
    return torch.complex(dt * sech2 * mr / rho + dth1 * (-mi / q1), dt * sech2 * mi / rho + dth1 * (a / q1))


def polar_roundtrip_backward(s, dout, eps=O.HPARAMS["atan2_eps"]):
    """Backward of the polar split + recombination in front of the iSTFT (network_functions.py:398-401 with 141-143):
    out = |s| (cos phi, sin phi), phi = atan2(Im s, Re s + eps).  Element-wise; identity up to O(eps)."""
    sr, si, gr, gi = s.real, s.imag, dout.real, dout.imag
    rho = torch.sqrt(sr * sr + si * si)
    a = sr + eps
    phi = torch.atan2(si, a)
    c, sn = torch.cos(phi), torch.sin(phi)
    drho = gr * c + gi * sn
    dphi = rho * (-gr * sn + gi * c)
    q = a * a + si * si
    return torch.complex(drho * sr / rho + dphi * (-si / q), drho * si / rho + dphi * (a / q))


def mask_tail_backward(raw, noisy_spec, g_clean_wave, g_noise_wave=None, eps=O.HPARAMS["atan2_eps"]):
    """The training step's first backward stage in one piece (the adjoint of the fused decoder[6] mask tail, variant dcs when
    `g_noise_wave` is given, dc otherwise): waveform gradients -> gradient w.r.t. the un-bounded decoder output `raw` (B,256,T).
    iSTFT adjoint -> polar adjoint -> combine adjoint (Ŝ = Y - Y·M, N̂ = Y·M; dM = conj(Y)·dN̂) -> bound_cRM adjoint twice."""
    T = raw.shape[-1]
    m1 = O.bound_crm(raw, eps)
    m2 = O.bound_crm(m1, eps)
    prod = _mul(noisy_spec, m2)
    if g_noise_wave is not None:
        clean = noisy_spec - prod
        dprod = polar_roundtrip_backward(prod, istft_adjoint(g_noise_wave, T), eps) \
            - polar_roundtrip_backward(clean, istft_adjoint(g_clean_wave, T), eps)
    else:
        dprod = polar_roundtrip_backward(prod, istft_adjoint(g_clean_wave, T), eps)
    dm2 = _mul(torch.conj(noisy_spec), dprod)
    return bound_crm_backward(raw, bound_crm_backward(m1, dm2, eps), eps)


def si_snr_backward(clean, estimate, eps=1e-8):
    """d SiSNR(clean, estimate) / d estimate (network_functions.py:30-42; mean over the batch rows), closed form: three
    row reductions (<e,c>, <c,c>, then |e - s_t|^2) and one element-wise pass — the loss kernel's contract."""
    rows = estimate.shape[0] if estimate.dim() > 1 else 1
    dot = torch.sum(estimate * clean, -1, keepdim=True)
    cc = torch.sum(clean * clean, -1, keepdim=True)
    k = dot / (cc + eps)
    s_t = k * clean
    e = estimate - s_t
    Tn = torch.sum(s_t * s_t, -1, keepdim=True)
    Nn = torch.sum(e * e, -1, keepdim=True)
    ratio = Tn / (Nn + eps) + eps
    dTn = 2 * k * cc / (cc + eps) * clean                      # d|s_t|^2 / d estimate
    dNn = 2 * (e - torch.sum(e * clean, -1, keepdim=True) / (cc + eps) * clean)
    return (10.0 / math.log(10.0)) / rows / ratio * (dTn / (Nn + eps) - Tn * dNn / (Nn + eps) ** 2)


def loss_backward(clean_audio, predict_clean_audio, noise_audio=None, predict_noise_audio=None, speech_alpha=0.7):
    """Waveform gradients of calc_loss with the config.py defaults (noise_loss_type 6, speech_loss_type 0):
    total = [1 - alpha * (-SiSNR(noise, n_hat))] + alpha * (-SiSNR(clean, s_hat))  (network_functions.py:195-204, precedence as
    written).  Returns (d total / d s_hat, d total / d n_hat or None)."""
    g_clean = -speech_alpha * si_snr_backward(clean_audio, predict_clean_audio)
    g_noise = None if noise_audio is None else speech_alpha * si_snr_backward(noise_audio, predict_noise_audio)
    return g_clean, g_noise


def cconv2d_backward(x, w_r, w_i, dy, stride, padding):
    """Backward of ComplexConv2d (`apply_complex`) in the PACKED real formulation the forward kernels
    use (one real implicit GEMM with N = 2 Cout, K = 2 Cin k^2): X = [x_re ; x_im], Wp = [[w_r, -w_i], [w_i, w_r]].
      dgrad : dX  = conv_transpose(dY, Wp)                    -> one GEMM, K = 2 Cout k^2
      wgrad : dWp = correlation(X, dY) (K = pixels)           -> one GEMM; its four blocks fold into
              dw_r = dWp[re,re] + dWp[im,im],  dw_i = dWp[im,re] - dWp[re,im]   (a 2-add epilogue)
      bias  : conv_r / conv_i each carry a bias (effective complex bias (b_r - b_i) + j (b_r + b_i)):
              db_r = sum(dY_re + dY_im), db_i = sum(dY_im - dY_re)
    Returns (dx complex, dw_r, dw_i, db_r, db_i)."""
    cout, cin = w_r.shape[0], w_r.shape[1]
    X = torch.cat([x.real, x.imag], dim=1)
    dY = torch.cat([dy.real, dy.imag], dim=1)
    Wp = torch.cat([torch.cat([w_r, -w_i], dim=1), torch.cat([w_i, w_r], dim=1)], dim=0)
    dX = torch.nn.grad.conv2d_input(X.shape, Wp, dY, stride=stride, padding=padding)
    dWp = torch.nn.grad.conv2d_weight(X, Wp.shape, dY, stride=stride, padding=padding)
    dw_r = dWp[:cout, :cin] + dWp[cout:, cin:]
    dw_i = dWp[cout:, :cin] - dWp[:cout, cin:]
    s = dY.sum(dim=[0, 2, 3])
    return torch.complex(dX[:, :cin], dX[:, cin:]), dw_r, dw_i, s[:cout] + s[cout:], s[cout:] - s[:cout]


def lstm_forward_saved(pre, whh):
    """One direction of one layer, forward in index order (reverse direction = flip the time axis outside): pre (B,S,4H) =
    x W_ih^T + b_ih + b_hh, gate order i,f,g,o.  Returns h (B,S,H) and what BPTT needs (activated gates, cell states)."""
    B, S, G = pre.shape
    H = G // 4
    h, c = pre.new_zeros(B, H), pre.new_zeros(B, H)
    hs, cs, gates = [], [], []
    for t in range(S):
        a = pre[:, t] + h @ whh.t()
        i, f, g, o = torch.sigmoid(a[:, :H]), torch.sigmoid(a[:, H:2 * H]), torch.tanh(a[:, 2 * H:3 * H]), torch.sigmoid(a[:, 3 * H:])
        c = f * c + i * g
        h = o * torch.tanh(c)
        hs.append(h), cs.append(c), gates.append(torch.cat([i, f, g, o], dim=1))
    return torch.stack(hs, 1), torch.stack(cs, 1), torch.stack(gates, 1)


def lstm_bptt(whh, hs, cs, gates, dh_out):
    """Back-propagation through time for lstm_forward_saved — the reverse-time recurrence the BPTT kernel runs (same shape as
    the forward recurrence: one H x 4H mat-vec per step, here with W_hh instead of W_hh^T).  Returns (dpre (B,S,4H), dW_hh);
    dW_ih = dpre^T x, db = sum dpre and dx = dpre W_ih are plain GEMMs / reductions outside the recurrence."""
    B, S, H = hs.shape
    dh_next, dc_next = hs.new_zeros(B, H), hs.new_zeros(B, H)
    dpre = hs.new_zeros(B, S, 4 * H)
    dW = torch.zeros_like(whh)
    for t in range(S - 1, -1, -1):
        i, f, g, o = gates[:, t, :H], gates[:, t, H:2 * H], gates[:, t, 2 * H:3 * H], gates[:, t, 3 * H:]
        tc = torch.tanh(cs[:, t])
        c_prev = cs[:, t - 1] if t > 0 else torch.zeros_like(tc)
        h_prev = hs[:, t - 1] if t > 0 else torch.zeros_like(tc)
        dh = dh_out[:, t] + dh_next
        dc = dh * o * (1 - tc * tc) + dc_next
        da = torch.cat([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g), dh * tc * o * (1 - o)], dim=1)
        dpre[:, t] = da
        dW += da.t() @ h_prev
        dh_next = da @ whh
        dc_next = dc * f
    return dpre, dW


def _cmul_conj(a, g):
    """dL/db for y = a * b (complex) in torch's gradient convention: conj(a) * g."""
    return _mul(torch.conj(a), g)


def attention_backward(x, sd, p_chan, p_spat, dy, k=7):
    """Backward of the attended product y = s * (a * x) (c_network.py:208-211 / 219-220): a = ComplexChannelAttention(x) =
    sigma_c(2 fc(avg x)) (the "max" pool is an average, network_functions.py:135-138), u = a * x,
    s = ComplexSpatialAttention(u) = sigma_c(conv7x7([mean_c u, max_c Re u + j max_c Im u])).
    Returns (dx, {parameter name: gradient}).  Structure for the kernels: pass 1 over (u, dy) reduces ds = sum_c conj(u) dy per
    pixel; a 7x7 transposed conv of the 1-channel gate gradient gives (dmean, dmax) per pixel; pass 2 forms
    du = conj(s) dy + dmean / C + [c == argmax] dmax, reduces da = sum_hw conj(x) du per channel and writes dx = conj(a) du;
    the gate MLP's backward adds a per-channel constant to dx."""
    B, C, H, W = x.shape
    ones = lambda t: t * (1 - t)  # noqa: E731
    # ---- forward, keeping what backward needs
    avg = torch.complex(x.real.mean([2, 3], keepdim=True), x.imag.mean([2, 3], keepdim=True))
    w1r, w1i = sd[p_chan + "fc.0.conv_r.weight"], sd[p_chan + "fc.0.conv_i.weight"]
    w2r, w2i = sd[p_chan + "fc.2.conv_r.weight"], sd[p_chan + "fc.2.conv_i.weight"]
    hid_pre = O.apply_complex(lambda t: torch.nn.functional.conv2d(t, w1r), lambda t: torch.nn.functional.conv2d(t, w1i), avg)
    hid = O.crelu(hid_pre)
    fc = O.apply_complex(lambda t: torch.nn.functional.conv2d(t, w2r), lambda t: torch.nn.functional.conv2d(t, w2i), hid)
    a = O.csigmoid(2 * fc)
    u = _mul(a, x)
    mean = torch.mean(u, dim=1, keepdim=True)
    mre, are = torch.max(u.real, dim=1, keepdim=True)
    mim, aim = torch.max(u.imag, dim=1, keepdim=True)
    stats = torch.cat([mean, torch.complex(mre, mim)], dim=1)
    w7r, w7i = sd[p_spat + "conv1.conv_r.weight"], sd[p_spat + "conv1.conv_i.weight"]
    s = O.csigmoid(O.apply_complex(lambda t: torch.nn.functional.conv2d(t, w7r, padding=k // 2),
                                   lambda t: torch.nn.functional.conv2d(t, w7i, padding=k // 2), stats))
    # ---- backward
    ds = _cmul_conj(u, dy).sum(dim=1, keepdim=True)
    dspre = torch.complex(ds.real * ones(s.real), ds.imag * ones(s.imag))
    dstats, d7r, d7i, _, _ = cconv2d_backward(stats, w7r, w7i, dspre, 1, k // 2)
    du = _cmul_conj(s, dy) + dstats[:, :1] / C
    dmax = dstats[:, 1:2]
    du = torch.complex(du.real + torch.zeros_like(u.real).scatter_(1, are, dmax.real),
                       du.imag + torch.zeros_like(u.imag).scatter_(1, aim, dmax.imag))
    da = _cmul_conj(x, du).sum(dim=[2, 3], keepdim=True)
    dfc = 2 * torch.complex(da.real * ones(a.real), da.imag * ones(a.imag))
    dhid, d2r, d2i, _, _ = cconv2d_backward(hid, w2r, w2i, dfc, 1, 0)
    dhid_pre = torch.complex(dhid.real * (hid_pre.real > 0), dhid.imag * (hid_pre.imag > 0))
    davg, d1r, d1i, _, _ = cconv2d_backward(avg, w1r, w1i, dhid_pre, 1, 0)
    dx = _cmul_conj(a, du) + davg / (H * W)
    grads = {p_spat + "conv1.conv_r.weight": d7r, p_spat + "conv1.conv_i.weight": d7i, p_chan + "fc.2.conv_r.weight": d2r,
             p_chan + "fc.2.conv_i.weight": d2i, p_chan + "fc.0.conv_r.weight": d1r, p_chan + "fc.0.conv_i.weight": d1i}
    return dx, grads


def decoder_stage_backward(d, skip, w_r, w_i, dy, up):
    """Backward of one decoder stage's linear part (c_network.py:214-216): z = cat(d, skip) -> nearest up-sampling by `up`
    -> ComplexConvTranspose2d(k3, s1, p1) with weights (Cin, Cout, 3, 3), in the packed formulation
    Wp[(in part), (out part)] = [[w_r, w_i], [-w_i, w_r]]:
      dgrad : dZup = conv2d(dY, Wp, padding 1)  (the adjoint of a stride-1 transposed conv is a plain conv with the same tensor)
      wgrad : dWp  = correlation(dY, Zup);  dw_r = dWp[re,re] + dWp[im,im],  dw_i = dWp[re,im] - dWp[im,re]
      up-sampling adjoint: sum over each uh x uw pixel block (an epilogue of the dgrad GEMM); concat adjoint: channel split.
    Returns (dd, dskip, dw_r, dw_i, db_r, db_i)."""
    F_ = torch.nn.functional
    cin, cout = w_r.shape[0], w_r.shape[1]
    z = torch.cat([d, skip], dim=1)
    zup = O.cupsample_nearest(z, up)
    Z = torch.cat([zup.real, zup.imag], dim=1)
    dY = torch.cat([dy.real, dy.imag], dim=1)
    Wp = torch.cat([torch.cat([w_r, w_i], dim=1), torch.cat([-w_i, w_r], dim=1)], dim=0)        # (2 Cin, 2 Cout, 3, 3)
    dZ = F_.conv2d(dY, Wp, padding=1)
    dWp = torch.nn.grad.conv2d_weight(dY, Wp.shape, Z, padding=1)
    dw_r = dWp[:cin, :cout] + dWp[cin:, cout:]
    dw_i = dWp[:cin, cout:] - dWp[cin:, :cout]
    B, _, Hu, Wu = dZ.shape
    dz = dZ.reshape(B, 2 * cin, Hu // up[0], up[0], Wu // up[1], up[1]).sum(dim=(3, 5))
    dz = torch.complex(dz[:, :cin], dz[:, cin:])
    s = dY.sum(dim=[0, 2, 3])
    cd = d.shape[1]
    return dz[:, :cd], dz[:, cd:], dw_r, dw_i, s[:cout] + s[cout:], s[cout:] - s[:cout]


def clinear_backward(x, w_r, w_i, dy):
    """ComplexLinear (c_network.py:124-126, 202) backward = the 1x1 case of cconv2d_backward on (B*S, in) rows.
    x, dy: (B, S, features) complex.  Returns (dx, dw_r, dw_i, db_r, db_i)."""
    B, S, _ = x.shape
    as_img = lambda t: t.reshape(B * S, -1, 1, 1)                        # noqa: E731
    dx, dwr, dwi, dbr, dbi = cconv2d_backward(as_img(x), w_r[:, :, None, None], w_i[:, :, None, None], as_img(dy), 1, 0)
    return dx.reshape(B, S, -1), dwr[:, :, 0, 0], dwi[:, :, 0, 0], dbr, dbi
