"""Evaluation step (`python test.py <variant>`'s per-batch function, network_functions.py:363-448):
`calc_loss`, SiSNR, wSDR and `test_batch_2_metric_loss` of the drop-in against vectors produced by EXECUTING the
reference's own functions (oracle/make_golden_eval.py -> tests/golden/eval_step.pt).

CPU: the loss code (plain tensor expressions, device-agnostic) on tensors rebuilt with the oracle.
GPU: the whole per-batch function through the C ABI kernels, all four variants.  (File sorts last on purpose: the parity
tests of the hot path run first.)"""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402
from oracle.make_golden_rnet import randomise_bn  # noqa: E402
from conftest import load_golden, rel_err, build_product_net  # noqa: E402

LOSS_TOL = 2e-3      # dB-scale losses; the fp32 kernel sequence is <= 1e-5 on the audio


def _batch(g):
    clean, noise, noisy = O.synthetic_audio(g["B"], 32 * (g["T"] - 1), seed=g["audio_seed"])
    return O.stft(noise), O.stft(noisy), O.stft(clean)


def test_sisnr_wsdr_match_reference_values():
    from dcsnet_b200 import network_functions as NF
    v = load_golden("eval_step.pt")["loss_vectors"]
    assert abs(float(NF.SiSNR()(v["a"], v["b"])) - v["sisnr"]) <= 1e-5 * max(1.0, abs(v["sisnr"]))
    assert abs(float(NF.wSDR()(v["a"], v["b"], v["c"])) - v["wsdr"]) <= 1e-6
    # closed form, independent of torch reductions: SI-SNR of a scaled copy plus orthogonal noise
    t = torch.linspace(0, 50, 4000)
    s, n = torch.sin(t)[None], torch.cos(t)[None]
    n = n - (n * s).sum() / (s * s).sum() * s
    want = 10 * torch.log10((s * s).sum() / (0.01 * (n * n).sum()))
    assert abs(float(NF.SiSNR()(s, 3.0 * s + 0.3 * n)) - float(10 * torch.log10(9 * (s * s).sum() / (0.09 * (n * n).sum())))) < 1e-3
    assert abs(float(NF.SiSNR()(s, s + 0.1 * n)) - float(want)) < 1e-3


def test_calc_loss_every_noise_loss_type_matches_reference():
    """calc_loss (network_functions.py:168-208) for noise_loss_type 0..6, on tensors rebuilt by the CPU oracle from the same
    seeds, equals what the reference's calc_loss returned on the reference's own tensors."""
    from dcsnet_b200 import network_functions as NF, config as C
    g = load_golden("eval_step.pt")
    net = build_product_net("randbn")
    sd = net.state_dict()
    nb, yb, cb = _batch(g)
    r = O.enhance_spec(sd, yb, "dcs")
    wave = lambda s: O.spec_to_wave(s)                                   # noqa: E731
    kw = dict(target_noise_mask=O.bound_crm(_crm(nb, yb)), predict_noise_mask=r["mask"],
              predict_noise_audio=wave(r["noise_spec"]), predict_clean_audio=wave(r["clean_spec"]),
              noise_audio=wave(nb), noisy_audio=wave(yb), clean_audio=wave(cb))
    for t, want in g["dcs_calc_loss_by_type"].items():
        hp = dict(C.hparams)
        hp["noise_loss_type"] = t
        fake = types.SimpleNamespace(hparams=hp, config=C.config)
        got = [float(x) for x in NF.calc_loss(fake, variant="dcs", **kw)]
        assert max(abs(a - b) for a, b in zip(got, want)) <= LOSS_TOL, (t, got, want)
    hp = dict(C.hparams)
    fake = types.SimpleNamespace(hparams=hp, config=C.config)
    assert abs(float(NF.calc_loss(fake, variant="dc", predict_clean_audio=kw["predict_clean_audio"], clean_audio=kw["clean_audio"]))
               - g["dcs_calc_loss_by_type"][6][1]) <= LOSS_TOL       # the speech term is the same expression
    with pytest.raises(ValueError):
        NF.calc_loss(fake, variant="xyz", **kw)


def _crm(S, Y, eps=1e-8):
    """network_functions.py:62-75."""
    den = Y.real ** 2 + Y.imag ** 2 + eps
    return torch.complex((Y.real * S.real + Y.imag * S.imag) / den, (Y.real * S.imag - Y.imag * S.real) / den)


def test_evaluation_step_refuses_cpu_tensors():
    from dcsnet_b200 import network_functions as NF
    g = load_golden("eval_step.pt")
    net = build_product_net("randbn")
    nb, yb, cb = _batch(g)
    with pytest.raises(RuntimeError):
        NF.test_batch_2_metric_loss(net, (nb, yb, cb, ["a", "b"], torch.zeros(2)), 0, "complex", variant="dcs")


def _install_stand_ins(monkeypatch):
    """Replace every kernel wrapper the evaluation step calls by the oracle's CPU function of the same contract."""
    from dcsnet_b200 import network_functions as NF, ops, r_network, config as C
    from oracle import rnet_oracle as RO
    monkeypatch.setattr(ops, "istft", lambda spec, audio=None, atan2_eps=1e-6, exact_polar=False: O.spec_to_wave(spec, atan2_eps))
    monkeypatch.setattr(ops, "mag_phase", lambda spec, eps=10e-7, want_phase=True: (torch.abs(spec), None))
    monkeypatch.setattr(NF, "cRM", _crm)
    monkeypatch.setattr(NF, "bound_cRM", lambda m, hp: O.bound_crm(m, hp["atan2_eps"]))

    def fake_complex(net, noisy, variant="dcs"):
        r = O.enhance_spec(net.state_dict(), noisy, variant)
        d = dict(predict_noise_mask=r["mask"], predict_clean_audio=O.spec_to_wave(r["clean_spec"]))
        if r["noise_spec"] is not None:
            d["predict_noise_audio"] = O.spec_to_wave(r["noise_spec"])
        return d

    def fake_real(net, noisy, variant="drs", atan2_eps=10e-7):
        r = RO.enhance_spec(net.state_dict(), noisy, variant)
        phase = torch.atan2(noisy.imag, noisy.real + atan2_eps)
        d = dict(predict_noise_mask=r["mask"], predict_clean_audio=r["clean_audio"], noisy_mag=torch.abs(noisy))
        if r["noise_mag"] is not None:
            d["predict_noise_audio"] = RO.mag_phase_2_wave(r["noise_mag"], phase)
        return d
    monkeypatch.setattr(NF, "enhance_batch", fake_complex)
    monkeypatch.setattr(r_network, "enhance_batch_real", fake_real)


@pytest.mark.parametrize("variant", ["dcs", "dc", "drs", "dr"])
def test_evaluation_step_plumbing_with_oracle_stand_ins(variant, monkeypatch):
    """Host logic only: with every kernel wrapper the step calls replaced by the oracle's CPU function of the same contract,
    val_ / test_batch_2_metric_loss must reproduce the reference's losses, audio and return-tuple layout for all variants.
    (The kernels themselves are checked by the -m gpu tests; this one pins the glue between them.)"""
    from dcsnet_b200 import network_functions as NF, r_network, config as C
    g = load_golden("eval_step.pt")
    want = g[variant]
    _install_stand_ins(monkeypatch)
    if variant in ("dcs", "dc"):
        net, dtype = build_product_net("randbn"), "complex"
    else:
        net = r_network.R_NETWORK(C.Config(), dict(C.hparams), 0).eval()
        randomise_bn(net.state_dict(), g["bn_seed"])
        dtype = "real"
    nb, yb, cb = _batch(g)
    r = NF.test_batch_2_metric_loss(net, (nb, yb, cb, ["id0", "id1"], torch.tensor([0, 0])), 0, dtype, variant=variant,
                                    metrics=dict(pesq=lambda c, p, sr: 1.0, stoi=lambda c, p, sr: float("nan")))
    assert len(r) == want["n_returned"]
    if variant in ("dcs", "drs"):
        assert abs(float(r[0]) - want["noise_loss"]) <= LOSS_TOL and abs(float(r[2]) - want["test_loss"]) <= LOSS_TOL
        assert abs(float(r[1]) - want["speech_loss"]) <= LOSS_TOL
        assert rel_err(r[5], want["predict_noise_audio"]) <= 1e-4 and rel_err(r[6], want["predict_clean_audio"]) <= 1e-4
        assert (r[3], r[4]) == (1.0, 0.0) and r[10] == ["id0", "id1"]          # NaN metric values are dropped (line 161)
    else:
        assert abs(float(r[0]) - want["speech_loss"]) <= LOSS_TOL and rel_err(r[3], want["predict_clean_audio"]) <= 1e-4
    rv = NF.val_batch_2_metric_loss(net, (nb, yb, cb, ["id0", "id1"]), 0, dtype, variant=variant, metrics=dict(pesq=lambda c, p, sr: 1.0))
    assert len(rv) == (10 if variant in ("dcs", "drs") else 7) and float(rv[0]) == float(r[0])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["dcs", "dc", "drs", "dr"])
def test_gpu_evaluation_step_matches_reference(variant):
    from dcsnet_b200 import network_functions as NF, r_network, config as C
    import dcsnet_b200 as D
    g = load_golden("eval_step.pt")
    want = g[variant]
    if variant in ("dcs", "dc"):
        net, dtype = build_product_net("randbn").cuda().eval(), "complex"
    else:
        net = r_network.R_NETWORK(C.Config(), dict(C.hparams), 0).eval()
        randomise_bn(net.state_dict(), g["bn_seed"])
        net, dtype = net.cuda().eval(), "real"
    nb, yb, cb = (t.cuda() for t in _batch(g))
    n0 = D._lib.launch_count()
    r = NF.test_batch_2_metric_loss(net, (nb, yb, cb, ["id0", "id1"], torch.tensor([0, 0])), 0, dtype, variant=variant,
                                    metrics=dict(pesq=lambda c, p, sr: 1.0, stoi=lambda c, p, sr: 0.5))
    torch.cuda.synchronize()
    assert D._lib.launch_count() - n0 >= 4 and len(r) == want["n_returned"]
    if variant in ("dcs", "drs"):
        noise_loss, speech_loss, test_loss, pesq_av, stoi_av, predict_noise_audio, predict_clean_audio = r[:7]
        assert abs(float(noise_loss) - want["noise_loss"]) <= LOSS_TOL and abs(float(test_loss) - want["test_loss"]) <= LOSS_TOL
        assert rel_err(predict_noise_audio, want["predict_noise_audio"]) <= 1e-4
        assert r[10] == ["id0", "id1"]
    else:
        speech_loss, pesq_av, stoi_av, predict_clean_audio = r[:4]
    assert abs(float(speech_loss) - want["speech_loss"]) <= LOSS_TOL
    assert rel_err(predict_clean_audio, want["predict_clean_audio"]) <= 1e-4
    assert (pesq_av, stoi_av) == (1.0, 0.5)
    # the validation function is the same step on a 4-tuple batch, without id / start_point (network_functions.py:282-361)
    rv = NF.val_batch_2_metric_loss(net, (nb, yb, cb, ["id0", "id1"]), 0, dtype, variant=variant)
    assert len(rv) == (10 if variant in ("dcs", "drs") else 7) and abs(float(rv[0]) - float(r[0])) <= 1e-6


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


@pytest.mark.parametrize("variant", ["dcs", "dr"])
def test_lightning_style_steps_with_oracle_stand_ins(variant, monkeypatch):
    """validation_step / test_step / configure_optimizers of the product networks (c_network.py:228-239, 263-302, 337-372):
    metric names, audio dictionary and NaN handling as in the reference; kernels replaced by oracle stand-ins (host logic)."""
    from dcsnet_b200 import r_network, config as C
    g = load_golden("eval_step.pt")
    _install_stand_ins(monkeypatch)
    if variant == "dcs":
        net = build_product_net("randbn")
    else:
        net = r_network.R_NETWORK(C.Config(), dict(C.hparams), 0).eval()
        randomise_bn(net.state_dict(), g["bn_seed"])
    net.variant = variant
    nb, yb, cb = _batch(g)
    out, metrics = net.test_step((nb, yb, cb, ["id0", "id1"], torch.tensor([0, 0])), 0)
    vout, vmetrics = net.validation_step((nb, yb, cb, ["id0", "id1"]), 0)
    if variant == "dcs":
        assert list(metrics) == ["test_loss", "test_noise_loss", "test_speech_loss", "test_pesq", "test_stoi"]
        assert list(out) == ["clean", "predict_clean", "noise", "predict_noise", "noisy"]
        assert abs(float(metrics["test_loss"]) - g["dcs"]["test_loss"]) <= LOSS_TOL
        assert list(vmetrics) == ["val_loss", "val_noise_loss", "val_speech_loss", "val_pesq", "val_stoi"]
        assert abs(float(vmetrics["val_loss"]) - g["dcs"]["test_loss"]) <= LOSS_TOL
    else:
        assert list(metrics) == ["test_speech_loss", "test_pesq", "test_stoi"] and list(out) == ["clean", "predict_clean", "noise", "noisy"]
        assert abs(float(metrics["test_speech_loss"]) - g["dr"]["speech_loss"]) <= LOSS_TOL
        assert list(vmetrics) == ["val_speech_loss", "val_pesq", "val_stoi"]
    assert out["predict_clean"].shape == (2, 32 * (g["T"] - 1)) and vout["clean"].dtype.name == "float32"
    opt = net.configure_optimizers()
    assert opt["monitor"] == ("val_loss" if variant == "dcs" else "speech_loss")
    assert opt["optimizer"].defaults["amsgrad"] is True and opt["optimizer"].defaults["lr"] == C.hparams["lr"]
    with pytest.raises(NotImplementedError):
        net.training_step(None, 0)


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
