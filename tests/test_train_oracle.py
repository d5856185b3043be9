"""Training step — ORACLE ONLY, no product kernels yet.  oracle/train_oracle.py restates the reference's
`train_batch_2_loss` + backward (train-mode batch norm, dropout, losses) and the closed-form adjoints the backward kernels
will implement; pins: tests/golden/train_step.pt (the reference's own losses / gradients / running statistics,
oracle/make_golden_train.py) and torch autograd.  All CPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dcsnet_oracle as O  # noqa: E402
from conftest import load_golden, rel_err, build_product_net  # noqa: E402


@pytest.mark.parametrize("variant", ["dcs", "drs"])
def test_training_step_fixture_lines_up_with_the_product_parameters(variant):
    """tests/golden/train_step.pt (reference `train_batch_2_loss` + backward, oracle/make_golden_train.py) is the pin for the
    training step that is still to be built.  Until then: the fixture must describe exactly the product
    containers' parameters (names and shapes), with finite losses and gradients, so the backward kernels have a target."""
    import math
    from dcsnet_b200 import r_network, config as C
    g = load_golden("train_step.pt")[variant]
    net = build_product_net("default") if variant == "dcs" else r_network.R_NETWORK(C.Config(), dict(C.hparams), 0)
    params = {k: tuple(torch.view_as_real(p).shape) if p.is_complex() else tuple(p.shape) for k, p in net.named_parameters()}
    assert set(params) == set(g["grads"]) | set(g["no_grad"])
    for k, f in g["grads"].items():
        assert params[k] == f["shape"], k
        assert math.isfinite(f["norm"]) and bool(torch.isfinite(f["head"]).all()), k
    assert all(math.isfinite(g[k]) for k in ("noise_loss", "speech_loss", "train_loss", "grad_norm"))
    assert abs(g["noise_loss"] + g["speech_loss"] - g["train_loss"]) < 1e-4
    assert abs(math.sqrt(sum(f["norm"] ** 2 for f in g["grads"].values())) - g["grad_norm"]) <= 1e-3 * g["grad_norm"]
    assert {k for k in net.state_dict() if "running_" in k} == set(g["running_stats"])


@pytest.mark.parametrize("variant", ["dcs", "drs"])
def test_training_step_oracle_matches_reference_gradients(variant):
    """oracle/train_oracle.py (train-mode restatement + autograd) against the reference's own `train_batch_2_loss` +
    `backward()` (tests/golden/train_step.pt): losses, every parameter gradient, BN running statistics after the step, for
    the complex (dcs) and the real (drs) network.  Conv biases that feed a batch-statistic BN have a mathematically zero
    gradient (1e-8 rounding noise on both sides): they are compared on an absolute scale."""
    from oracle import train_oracle as TO
    from dcsnet_b200 import r_network, config as C
    g = load_golden("train_step.pt")
    w = g[variant]
    net = build_product_net("default") if variant == "dcs" else r_network.R_NETWORK(C.Config(), dict(C.hparams), 0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = {k for k, _ in net.named_parameters()}
    clean, noise, noisy = O.synthetic_audio(g["B"], 32 * (g["T"] - 1), seed=g["audio_seed"])
    step = TO.train_step if variant == "dcs" else TO.train_step_real
    r = step(sd, O.stft(noise), O.stft(noisy), O.stft(clean), names, variant)
    for k in ("noise_loss", "speech_loss", "train_loss"):
        assert abs(r[k] - w[k]) <= 1e-4, k
    assert set(r["grads"]) == set(w["grads"])
    floor = 1e-6 * w["grad_norm"]
    checked = 0
    for k, f in w["grads"].items():
        gr = r["grads"][k].float()
        assert abs(float(gr.norm()) - f["norm"]) <= 1e-3 * f["norm"] + floor, k
        assert float((gr.reshape(-1)[:8] - f["head"]).abs().max()) <= 1e-3 * f["max_abs"] + floor, k
        checked += f["norm"] > 100 * floor
    assert checked > 0.75 * len(w["grads"])                   # the bulk of the gradients is far above the noise floor
    n_stats = 0
    for k, f in w["running_stats"].items():
        if k.endswith("num_batches_tracked"):
            continue
        t = r["running_stats"][k]
        t = (torch.view_as_real(t) if t.is_complex() else t).float()
        assert abs(float(t.norm()) - f["norm"]) <= 1e-5 * f["norm"] and float((t.reshape(-1)[:8] - f["head"]).abs().max()) <= 1e-5 * f["max_abs"], k
        n_stats += 1
    assert n_stats == 2 * 14                                   # initial BN + 7 encoder + 6 decoder layers, two statistics each


def test_training_step_oracle_dropout_positions_match_reference():
    """With the reference's dropout probabilities (0.1 conv / 0.2 fc) and torch's CPU generator seeded as in the fixture run,
    the restatement draws the same masks in the same order (7 encoder outputs, fc output, 7 decoder outputs; real and
    imaginary parts independently) and must land on the reference's losses."""
    from oracle import train_oracle as TO
    g = load_golden("train_step.pt")
    w = g["dcs_dropout"]
    net = build_product_net("default")
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = {k for k, _ in net.named_parameters()}
    clean, noise, noisy = O.synthetic_audio(g["B"], 32 * (g["T"] - 1), seed=g["audio_seed"])
    specs = O.stft(noise), O.stft(noisy), O.stft(clean)
    torch.manual_seed(w["seed"])
    r = TO.train_step(sd, *specs, names, "dcs", drop=TO.dropout_torch_stream(w["dropout_conv"], w["dropout_fc"]))
    for k in ("noise_loss", "speech_loss", "train_loss"):
        assert abs(r[k] - w[k]) <= 1e-4, (k, r[k], w[k])
    assert abs(w["train_loss"] - g["dcs"]["train_loss"]) > 0.1          # the dropout case really differs from p = 0


@pytest.mark.parametrize("B,T", [(2, 40), (1, 17)])
def test_istft_adjoint_closed_form_equals_autograd(B, T):
    """oracle/train_oracle.istft_adjoint — the contract of the training step's first backward kernel (the loss is SI-SNR on
    waveforms, so every gradient enters through the iSTFT) — against torch autograd through the reference's mag_phase_2_wave."""
    from oracle import train_oracle as TO, rnet_oracle as RO
    gen = torch.Generator().manual_seed(B * 100 + T)
    mag = torch.rand(B, 256, T, generator=gen).requires_grad_(True)
    phase = (6.28 * torch.rand(B, 256, T, generator=gen) - 3.14).requires_grad_(True)
    y = RO.mag_phase_2_wave(mag, phase)
    g = torch.randn(y.shape, generator=gen)
    (y * g).sum().backward()
    gs = TO.istft_adjoint(g, T)                                         # gradient w.r.t. the complex spectrogram
    want_mag = gs.real * torch.cos(phase.detach()) + gs.imag * torch.sin(phase.detach())
    want_phase = mag.detach() * (-gs.real * torch.sin(phase.detach()) + gs.imag * torch.cos(phase.detach()))
    assert rel_err(want_mag, mag.grad) <= 1e-5 and rel_err(want_phase, phase.grad) <= 1e-5


@pytest.mark.parametrize("shape", [(3, 8, 6, 10), (2, 1, 16, 5), (4, 32, 2, 7)])
def test_complex_batchnorm_train_backward_closed_form_equals_autograd(shape):
    """oracle/train_oracle.cbn_train_backward (two reduction passes + a per-channel 3x3 Jacobian) against autograd through the
    train-mode restatement that itself reproduces the reference's gradients."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(sum(shape))
    Cn = shape[1]
    x = torch.complex(torch.randn(shape, generator=gen) * 1.5 + 0.3, torch.randn(shape, generator=gen) * 0.7 - 0.2)
    x = torch.complex(x.real, x.imag + 0.4 * x.real).requires_grad_(True)            # correlated parts: Cri != 0
    sd = {"p.weight": torch.stack([1 + 0.3 * torch.rand(Cn, generator=gen), 1 + 0.3 * torch.rand(Cn, generator=gen),
                                   0.3 * torch.rand(Cn, generator=gen) - 0.15], dim=1).requires_grad_(True),
          "p.bias": (0.1 * torch.randn(Cn, 2, generator=gen)).requires_grad_(True),
          "p.running_mean": torch.zeros(Cn, dtype=torch.complex64), "p.running_covar": torch.ones(Cn, 3)}
    y = TO.cbn_train({})(x, sd, "p.")
    dy = torch.complex(torch.randn(shape, generator=gen), torch.randn(shape, generator=gen))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    dx, dw, db = TO.cbn_train_backward(x.detach(), dy, sd["p.weight"].detach())
    assert rel_err(dx, x.grad) <= 2e-5 and rel_err(dw, sd["p.weight"].grad) <= 2e-5 and rel_err(db, sd["p.bias"].grad) <= 2e-5


def test_bound_crm_backward_closed_form_equals_autograd():
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(21)
    m = torch.complex(torch.randn(4, 256, 9, generator=gen) * 1.3, torch.randn(4, 256, 9, generator=gen) * 0.8).requires_grad_(True)
    out = O.bound_crm(O.bound_crm(m))                                   # the training path applies it twice
    dout = torch.complex(torch.randn(out.shape, generator=gen), torch.randn(out.shape, generator=gen))
    (out.real * dout.real + out.imag * dout.imag).sum().backward()
    inner = O.bound_crm(m.detach())
    got = TO.bound_crm_backward(m.detach(), TO.bound_crm_backward(inner, dout))
    assert rel_err(got, m.grad) <= 2e-5


@pytest.mark.parametrize("variant", ["dcs", "dc"])
def test_mask_tail_backward_closed_form_equals_autograd(variant):
    """Waveform gradients -> gradient at the decoder output through iSTFT, polar split, combine and the two bound_cRM, as one
    closed-form stage (the adjoint of the forward's fused mask tail) vs autograd through the oracle's forward functions."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(5)
    B, T = 2, 24
    raw = torch.complex(torch.randn(B, 256, T, generator=gen), torch.randn(B, 256, T, generator=gen)).requires_grad_(True)
    Y = O.stft(O.synthetic_audio(B, 32 * (T - 1))[2])
    m2 = O.bound_crm(O.bound_crm(raw))
    prod = torch.complex(Y.real * m2.real - Y.imag * m2.imag, Y.real * m2.imag + Y.imag * m2.real)
    gc = torch.randn(B, 32 * (T - 1), generator=gen)
    gn = torch.randn(B, 32 * (T - 1), generator=gen)
    if variant == "dcs":
        loss = (O.spec_to_wave(Y - prod) * gc).sum() + (O.spec_to_wave(prod) * gn).sum()
    else:
        loss = (O.spec_to_wave(prod) * gc).sum()
    loss.backward()
    got = TO.mask_tail_backward(raw.detach(), Y, gc, gn if variant == "dcs" else None)
    assert rel_err(got, raw.grad) <= 5e-5


def test_loss_backward_closed_form_and_full_tail_chain_equal_autograd():
    """calc_loss's waveform gradients in closed form (SI-SNR adjoint), then the whole chain loss -> waveforms -> mask tail ->
    decoder output against autograd through the product's calc_loss and the oracle's forward functions."""
    import types
    from oracle import train_oracle as TO
    from dcsnet_b200 import network_functions as NF, config as C
    gen = torch.Generator().manual_seed(17)
    B, T = 2, 24
    clean, noise, noisy = O.synthetic_audio(B, 32 * (T - 1))
    est = (clean + 0.3 * torch.randn(clean.shape, generator=gen)).requires_grad_(True)
    NF.SiSNR()(clean, est).backward()
    assert rel_err(TO.si_snr_backward(clean, est.detach()), est.grad) <= 1e-5
    raw = torch.complex(torch.randn(B, 256, T, generator=gen), torch.randn(B, 256, T, generator=gen)).requires_grad_(True)
    Y, Nn, S = O.stft(noisy), O.stft(noise), O.stft(clean)
    m2 = O.bound_crm(O.bound_crm(raw))
    prod = torch.complex(Y.real * m2.real - Y.imag * m2.imag, Y.real * m2.imag + Y.imag * m2.real)
    s_hat, n_hat = O.spec_to_wave(Y - prod), O.spec_to_wave(prod)
    clean_audio, noise_audio = O.spec_to_wave(S), O.spec_to_wave(Nn)
    fake = types.SimpleNamespace(hparams=dict(C.hparams), config=C.config)
    _, _, total = NF.calc_loss(fake, variant="dcs", predict_noise_audio=n_hat, predict_clean_audio=s_hat, noise_audio=noise_audio,
                               noisy_audio=O.spec_to_wave(Y), clean_audio=clean_audio, target_noise_mask=None, predict_noise_mask=None)
    total.backward()
    gc, gn = TO.loss_backward(clean_audio, s_hat.detach(), noise_audio, n_hat.detach(), C.hparams["speech_alpha"])
    got = TO.mask_tail_backward(raw.detach(), Y, gc, gn)
    assert rel_err(got, raw.grad) <= 1e-4


@pytest.mark.parametrize("cin,cout,k,stride", [(1, 8, 7, (2, 2)), (16, 32, 5, (2, 2)), (64, 128, 3, (2, 1))])
def test_complex_conv_backward_in_packed_formulation_equals_autograd(cin, cout, k, stride):
    """dgrad / wgrad / bias gradients of ComplexConv2d as ONE packed real GEMM each (the formulation the forward kernels
    already use) vs autograd through the oracle's four-real-convolution restatement, at encoder layer shapes."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(cin + cout)
    rnd = lambda *s: torch.randn(*s, generator=gen)                      # noqa: E731
    sd = {"c.conv_r.weight": (0.2 * rnd(cout, cin, k, k)).requires_grad_(True), "c.conv_i.weight": (0.2 * rnd(cout, cin, k, k)).requires_grad_(True),
          "c.conv_r.bias": rnd(cout).requires_grad_(True), "c.conv_i.bias": rnd(cout).requires_grad_(True)}
    x = torch.complex(rnd(2, cin, 16, 12), rnd(2, cin, 16, 12)).requires_grad_(True)
    y = O.cconv2d(x, sd, "c.", stride, k // 2)
    dy = torch.complex(rnd(*y.shape), rnd(*y.shape))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    dx, dwr, dwi, dbr, dbi = TO.cconv2d_backward(x.detach(), sd["c.conv_r.weight"].detach(), sd["c.conv_i.weight"].detach(), dy, stride, k // 2)
    for got, want in ((dx, x.grad), (dwr, sd["c.conv_r.weight"].grad), (dwi, sd["c.conv_i.weight"].grad),
                      (dbr, sd["c.conv_r.bias"].grad), (dbi, sd["c.conv_i.bias"].grad)):
        assert rel_err(got, want) <= 2e-5


@pytest.mark.parametrize("B,S,H", [(3, 8, 64), (2, 5, 128)])
def test_lstm_bptt_explicit_recurrence_equals_autograd(B, S, H):
    """The reverse-time recurrence of the BPTT kernel (ComplexLSTM hidden 64, real LSTM hidden 128) vs autograd."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(B + S + H)
    pre = (0.5 * torch.randn(B, S, 4 * H, generator=gen)).requires_grad_(True)
    whh = (0.1 * torch.randn(4 * H, H, generator=gen)).requires_grad_(True)
    hs, cs, gates = TO.lstm_forward_saved(pre, whh)
    dh = torch.randn(B, S, H, generator=gen)
    (hs * dh).sum().backward()
    dpre, dW = TO.lstm_bptt(whh.detach(), hs.detach(), cs.detach(), gates.detach(), dh)
    assert rel_err(dpre, pre.grad) <= 2e-5 and rel_err(dW, whh.grad) <= 2e-5
    # and the forward agrees with torch.nn.LSTM (what the reference calls, c_network.py:19-28)
    m = torch.nn.LSTM(7, H, batch_first=True)
    x = torch.randn(B, S, 7, generator=gen)
    with torch.no_grad():
        ref, _ = m(x)
        mine, _, _ = TO.lstm_forward_saved(x @ m.weight_ih_l0.t() + m.bias_ih_l0 + m.bias_hh_l0, m.weight_hh_l0)
    assert rel_err(mine, ref) <= 1e-5


@pytest.mark.parametrize("C,H,W", [(16, 12, 9), (128, 2, 11), (8, 32, 20)])
def test_attention_backward_closed_form_equals_autograd(C, H, W):
    """Channel + spatial attention and both complex products as one backward stage (two passes over the tensor, like the
    forward's streaming kernel) vs autograd through the oracle's forward functions."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(C + H + W)
    rnd = lambda *s: torch.randn(*s, generator=gen)                      # noqa: E731
    R = max(C // 16, 1)
    sd = {"a.fc.0.conv_r.weight": 0.3 * rnd(R, C, 1, 1), "a.fc.0.conv_i.weight": 0.3 * rnd(R, C, 1, 1),
          "a.fc.2.conv_r.weight": 0.5 * rnd(C, R, 1, 1), "a.fc.2.conv_i.weight": 0.5 * rnd(C, R, 1, 1),
          "s.conv1.conv_r.weight": 0.2 * rnd(1, 2, 7, 7), "s.conv1.conv_i.weight": 0.2 * rnd(1, 2, 7, 7)}
    sd = {k_: v.requires_grad_(True) for k_, v in sd.items()}
    x = torch.complex(rnd(2, C, H, W), rnd(2, C, H, W)).requires_grad_(True)
    u = O.channel_attention(x, sd, "a.") * x
    y = O.spatial_attention(u, sd, "s.") * u
    dy = torch.complex(rnd(*y.shape), rnd(*y.shape))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    dx, grads = TO.attention_backward(x.detach(), {k_: v.detach() for k_, v in sd.items()}, "a.", "s.", dy)
    assert rel_err(dx, x.grad) <= 5e-5
    for k_, g_ in grads.items():
        assert rel_err(g_, sd[k_].grad) <= 5e-5, k_


@pytest.mark.parametrize("cd,cs,cout,up", [(128, 128, 128, (2, 1)), (16, 16, 8, (2, 2)), (8, 8, 1, (2, 2))])
def test_decoder_stage_backward_in_packed_formulation_equals_autograd(cd, cs, cout, up):
    """concat + nearest up-sampling + ComplexConvTranspose2d backward (dgrad as a plain conv with a pixel-block-sum epilogue,
    wgrad with block folding) vs autograd through the oracle's forward functions, at decoder shapes incl. decoder[6]."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(cd + cout)
    rnd = lambda *s: torch.randn(*s, generator=gen)                      # noqa: E731
    cin = cd + cs
    sd = {"t.conv_tran_r.weight": (0.1 * rnd(cin, cout, 3, 3)).requires_grad_(True), "t.conv_tran_i.weight": (0.1 * rnd(cin, cout, 3, 3)).requires_grad_(True),
          "t.conv_tran_r.bias": rnd(cout).requires_grad_(True), "t.conv_tran_i.bias": rnd(cout).requires_grad_(True)}
    d = torch.complex(rnd(2, cd, 4, 6), rnd(2, cd, 4, 6)).requires_grad_(True)
    skip = torch.complex(rnd(2, cs, 4, 6), rnd(2, cs, 4, 6)).requires_grad_(True)
    y = O.cconvT2d(O.cupsample_nearest(torch.cat((d, skip), dim=1), up), sd, "t.", 1, 1)
    dy = torch.complex(rnd(*y.shape), rnd(*y.shape))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    got = TO.decoder_stage_backward(d.detach(), skip.detach(), sd["t.conv_tran_r.weight"].detach(), sd["t.conv_tran_i.weight"].detach(), dy, up)
    want = (d.grad, skip.grad, sd["t.conv_tran_r.weight"].grad, sd["t.conv_tran_i.weight"].grad, sd["t.conv_tran_r.bias"].grad, sd["t.conv_tran_i.bias"].grad)
    for g_, w_ in zip(got, want):
        assert rel_err(g_, w_) <= 2e-5


def test_encoder_layer_backward_chain_equals_autograd():
    """The contracts composed as an encoder layer runs them backward: ComplexReLU mask -> cbn_train_backward ->
    cconv2d_backward, vs autograd through conv + train-mode BN + ReLU.  (Also shows why the reference's conv biases get a
    zero gradient in train mode: dY into the conv has zero mean per channel after the BN backward.)"""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(3)
    rnd = lambda *s: torch.randn(*s, generator=gen)                      # noqa: E731
    cin, cout, k, stride = 8, 16, 5, (2, 2)
    sd = {"e.0.conv_r.weight": (0.2 * rnd(cout, cin, k, k)).requires_grad_(True), "e.0.conv_i.weight": (0.2 * rnd(cout, cin, k, k)).requires_grad_(True),
          "e.0.conv_r.bias": rnd(cout).requires_grad_(True), "e.0.conv_i.bias": rnd(cout).requires_grad_(True),
          "e.1.weight": torch.stack([1 + 0.2 * torch.rand(cout, generator=gen), 1 + 0.2 * torch.rand(cout, generator=gen),
                                     0.2 * torch.rand(cout, generator=gen) - 0.1], dim=1).requires_grad_(True),
          "e.1.bias": (0.1 * rnd(cout, 2)).requires_grad_(True),
          "e.1.running_mean": torch.zeros(cout, dtype=torch.complex64), "e.1.running_covar": torch.ones(cout, 3)}
    x = torch.complex(rnd(3, cin, 16, 12), rnd(3, cin, 16, 12)).requires_grad_(True)
    conv = O.cconv2d(x, sd, "e.0.", stride, k // 2)
    bn = TO.cbn_train({})(conv, sd, "e.1.")
    y = O.crelu(bn)
    dy = torch.complex(rnd(*y.shape), rnd(*y.shape))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    dbn = torch.complex(dy.real * (bn.real > 0), dy.imag * (bn.imag > 0)).detach()
    dconv, dw, db = TO.cbn_train_backward(conv.detach(), dbn, sd["e.1.weight"].detach())
    dx, dwr, dwi, dbr, dbi = TO.cconv2d_backward(x.detach(), sd["e.0.conv_r.weight"].detach(), sd["e.0.conv_i.weight"].detach(), dconv, stride, k // 2)
    assert rel_err(dx, x.grad) <= 5e-5 and rel_err(dwr, sd["e.0.conv_r.weight"].grad) <= 5e-5 and rel_err(dwi, sd["e.0.conv_i.weight"].grad) <= 5e-5
    assert rel_err(dw, sd["e.1.weight"].grad) <= 5e-5 and rel_err(db, sd["e.1.bias"].grad) <= 5e-5
    assert float(dbr.abs().max()) <= 1e-4 * float(dwr.abs().max()) and float(dbi.abs().max()) <= 1e-4 * float(dwr.abs().max())


def test_decoder_stage_full_backward_chain_equals_autograd():
    """A whole decoder stage backward from the contracts: attention_backward (decoder attention) -> LeakyReLU mask ->
    cbn_train_backward -> decoder_stage_backward (convT, up-sampling, concat), and the skip branch through attention_backward
    (skip attention), vs autograd through the oracle's forward functions (c_network.py:207-220)."""
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(8)
    rnd = lambda *s: torch.randn(*s, generator=gen)                      # noqa: E731
    C, cout, up = 32, 16, (2, 2)
    R_in, R_out = max(C // 16, 1), max(cout // 16, 1)
    sd = {"sk.a.fc.0.conv_r.weight": 0.3 * rnd(R_in, C, 1, 1), "sk.a.fc.0.conv_i.weight": 0.3 * rnd(R_in, C, 1, 1),
          "sk.a.fc.2.conv_r.weight": 0.5 * rnd(C, R_in, 1, 1), "sk.a.fc.2.conv_i.weight": 0.5 * rnd(C, R_in, 1, 1),
          "sk.s.conv1.conv_r.weight": 0.2 * rnd(1, 2, 7, 7), "sk.s.conv1.conv_i.weight": 0.2 * rnd(1, 2, 7, 7),
          "da.a.fc.0.conv_r.weight": 0.3 * rnd(R_out, cout, 1, 1), "da.a.fc.0.conv_i.weight": 0.3 * rnd(R_out, cout, 1, 1),
          "da.a.fc.2.conv_r.weight": 0.5 * rnd(cout, R_out, 1, 1), "da.a.fc.2.conv_i.weight": 0.5 * rnd(cout, R_out, 1, 1),
          "da.s.conv1.conv_r.weight": 0.2 * rnd(1, 2, 7, 7), "da.s.conv1.conv_i.weight": 0.2 * rnd(1, 2, 7, 7),
          "t.0.conv_tran_r.weight": 0.1 * rnd(2 * C, cout, 3, 3), "t.0.conv_tran_i.weight": 0.1 * rnd(2 * C, cout, 3, 3),
          "t.0.conv_tran_r.bias": rnd(cout), "t.0.conv_tran_i.bias": rnd(cout),
          "t.1.weight": torch.stack([1 + 0.2 * torch.rand(cout, generator=gen), 1 + 0.2 * torch.rand(cout, generator=gen),
                                     0.2 * torch.rand(cout, generator=gen) - 0.1], dim=1), "t.1.bias": 0.1 * rnd(cout, 2)}
    sd = {k_: v.requires_grad_(True) for k_, v in sd.items()}
    sd.update({"t.1.running_mean": torch.zeros(cout, dtype=torch.complex64), "t.1.running_covar": torch.ones(cout, 3)})
    d = torch.complex(rnd(2, C, 4, 6), rnd(2, C, 4, 6)).requires_grad_(True)
    skip = torch.complex(rnd(2, C, 4, 6), rnd(2, C, 4, 6)).requires_grad_(True)
    ca = O.channel_attention(skip, sd, "sk.a.") * skip
    sa = O.spatial_attention(ca, sd, "sk.s.") * ca
    lin = O.cconvT2d(O.cupsample_nearest(torch.cat((d, sa), dim=1), up), sd, "t.0.", 1, 1)
    bn = TO.cbn_train({})(lin, sd, "t.1.")
    act = O.clrelu(bn)
    y = act * O.channel_attention(act, sd, "da.a.")
    y = y * O.spatial_attention(y, sd, "da.s.")
    dy = torch.complex(rnd(*y.shape), rnd(*y.shape))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    det = {k_: v.detach() for k_, v in sd.items()}
    dact, g_da = TO.attention_backward(act.detach(), det, "da.a.", "da.s.", dy)
    slope = lambda t: torch.where(t > 0, torch.ones_like(t), torch.full_like(t, O.LRELU_SLOPE))   # noqa: E731
    dbn = torch.complex(dact.real * slope(bn.real.detach()), dact.imag * slope(bn.imag.detach()))
    dlin, dw_bn, db_bn = TO.cbn_train_backward(lin.detach(), dbn, det["t.1.weight"])
    dd, dsa, dwr, dwi, _, _ = TO.decoder_stage_backward(d.detach(), sa.detach(), det["t.0.conv_tran_r.weight"], det["t.0.conv_tran_i.weight"], dlin, up)
    dskip, g_sk = TO.attention_backward(skip.detach(), det, "sk.a.", "sk.s.", dsa)
    assert rel_err(dd, d.grad) <= 1e-4 and rel_err(dskip, skip.grad) <= 1e-4
    assert rel_err(dwr, sd["t.0.conv_tran_r.weight"].grad) <= 1e-4 and rel_err(dw_bn, sd["t.1.weight"].grad) <= 1e-4
    for k_, g_ in list(g_da.items()) + list(g_sk.items()):
        assert rel_err(g_, sd[k_].grad) <= 1e-4, k_


def test_complex_linear_backward_equals_autograd():
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(4)
    rnd = lambda *s: torch.randn(*s, generator=gen)                      # noqa: E731
    sd = {"fc.fc_r.weight": (0.1 * rnd(128, 128)).requires_grad_(True), "fc.fc_i.weight": (0.1 * rnd(128, 128)).requires_grad_(True),
          "fc.fc_r.bias": rnd(128).requires_grad_(True), "fc.fc_i.bias": rnd(128).requires_grad_(True)}
    x = torch.complex(rnd(2, 10, 128), rnd(2, 10, 128)).requires_grad_(True)
    y = O.clinear(x, sd, "fc.")
    dy = torch.complex(rnd(*y.shape), rnd(*y.shape))
    (y.real * dy.real + y.imag * dy.imag).sum().backward()
    got = TO.clinear_backward(x.detach(), sd["fc.fc_r.weight"].detach(), sd["fc.fc_i.weight"].detach(), dy)
    want = (x.grad, sd["fc.fc_r.weight"].grad, sd["fc.fc_i.weight"].grad, sd["fc.fc_r.bias"].grad, sd["fc.fc_i.bias"].grad)
    for g_, w_ in zip(got, want):
        assert rel_err(g_, w_) <= 2e-5
