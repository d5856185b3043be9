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
    if variant in ("dr", "drs"):
        with pytest.raises(NotImplementedError):                      # the real path's training step is not built
            net.training_step(None, 0)
    else:                                                             # the complex path trains on the GPU only: no CPU fallback
        z = torch.zeros(1, 256, 16, dtype=torch.complex64)
        with pytest.raises(RuntimeError, match="CUDA"):
            net.training_step((z, z, z, ["x"]), 0)
