"""Generate tests/golden/eval_step.pt by EXECUTING THE REFERENCE'S per-batch evaluation function (build container only;
TEST INFRASTRUCTURE).

    python -m oracle.make_golden_eval

`network_functions.test_batch_2_metric_loss` (network_functions.py:363-448) and `calc_loss` (168-208) are imported
unmodified behind oracle/stubs and called on seeded batches for all four variants (dcs / dc on C_NETWORK, drs / dr on
R_NETWORK, seed-0 weights, randomised BN state).  One runtime substitution: `mag_phase_2_wave` hard-codes a cuda window
(network_functions.py:147), so the module attribute is pointed at the device-agnostic twin in reference_harness for the
duration of the call.  PESQ / STOI are stubs (NaN): third-party CPU metrics.  Also records `calc_loss` for every
`noise_loss_type` 0..6 on the dcs tensors, and SiSNR / wSDR on a seeded pair.
"""
import importlib
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_harness as rh, dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402
from oracle.make_golden_rnet import randomise_bn  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B, T = 2, 64


def batch():
    clean, noise, noisy = O.synthetic_audio(B, 32 * (T - 1))
    return rh.reference_stft(noise), rh.reference_stft(noisy), rh.reference_stft(clean), ["id0", "id1"], torch.tensor([0, 0])


def main():
    mods = rh.load("dcs")
    nf = mods["network_functions"]
    orig = nf.mag_phase_2_wave
    nf.mag_phase_2_wave = rh.mag_phase_2_wave_cpu
    out = dict(B=B, T=T, audio_seed=1234, bn_seed=7)
    try:
        cnet = rh.build_c_network(0, randomise_bn=False)
        SW.randomise_bn_state(cnet.state_dict(), 7)
        cnet.eval()
        with rh.argv_variant("drs"):
            rn = importlib.import_module("r_network")
            rnet = rn.R_NETWORK(mods["config"].Config(), dict(mods["config"].hparams), 0).eval()
        randomise_bn(rnet.state_dict(), 7)
        for variant, net, dtype in (("dcs", cnet, "complex"), ("dc", cnet, "complex"), ("drs", rnet, "real"), ("dr", rnet, "real")):
            with torch.no_grad(), rh.argv_variant(variant):
                r = nf.test_batch_2_metric_loss(net, batch(), 0, dtype)
            if variant in ("dcs", "drs"):
                names = ["noise_loss", "speech_loss", "test_loss", "pesq_av", "stoi_av", "predict_noise_audio", "predict_clean_audio",
                         "noise_audio", "noisy_audio", "clean_audio", "id", "start_point"]
            else:
                names = ["speech_loss", "pesq_av", "stoi_av", "predict_clean_audio", "noise_audio", "noisy_audio", "clean_audio"]
            assert len(r) == len(names)
            d = dict(zip(names, r))
            out[variant] = {k: (float(v) if k.endswith("_loss") else v) for k, v in d.items()
                            if k.endswith("_loss") or k in ("predict_clean_audio", "predict_noise_audio")}
            out[variant]["n_returned"] = len(r)
            if variant == "dcs":
                kw = dict(predict_noise_audio=d["predict_noise_audio"], predict_clean_audio=d["predict_clean_audio"],
                          noise_audio=d["noise_audio"], noisy_audio=d["noisy_audio"], clean_audio=d["clean_audio"])
                nb, yb, _, _, _ = batch()
                with torch.no_grad():
                    kw["target_noise_mask"] = nf.bound_cRM(nf.cRM(nb, yb), cnet.hparams)
                    kw["predict_noise_mask"] = nf.bound_cRM(cnet(yb), cnet.hparams)
                by_type = {}
                for t in range(7):
                    hp = dict(cnet.hparams)
                    hp["noise_loss_type"] = t
                    fake = types.SimpleNamespace(hparams=hp, config=cnet.config)
                    with torch.no_grad(), rh.argv_variant("dcs"):
                        by_type[t] = [float(x) for x in nf.calc_loss(fake, **kw)]
                out["dcs_calc_loss_by_type"] = by_type
        g = torch.Generator().manual_seed(5)
        a, b, c = (torch.randn(3, 4000, generator=g) for _ in range(3))
        out["loss_vectors"] = dict(a=a, b=b, c=c, sisnr=float(nf.SiSNR()(a, b)), wsdr=float(nf.wSDR()(a, b, c)))
    finally:
        nf.mag_phase_2_wave = orig
    path = os.path.join(OUT, "eval_step.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: out[k].get("test_loss", out[k].get("speech_loss")) for k in ("dcs", "dc", "drs", "dr")})
    print(out["dcs_calc_loss_by_type"])


if __name__ == "__main__":
    main()
