"""Generate tests/golden/rnet_*.pt by EXECUTING THE REFERENCE'S r_network.py (build container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden_rnet

R_NETWORK(config, hparams, seed=0) from /root/reference/r_network.py, imported unmodified behind oracle/stubs, eval mode,
BatchNorm running statistics / affine parameters randomised from a seed (default BN state is the identity and would hide
folding bugs).  The fixture stores the state_dict key list with shapes, the sha256 of the seed-0 weights, the seeded
noisy spectrogram's magnitude mask, per-layer fingerprints and the dr / drs enhanced magnitudes / audio computed with the
reference's own step-function lines (network_functions.py:296-305, 338-342) and mag_phase_2_wave twin.
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_harness as rh, dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def randomise_bn(sd, seed):
    g = torch.Generator().manual_seed(seed)
    for k in list(sd):
        if k.endswith("running_mean"):
            sd[k].copy_(0.2 * torch.randn(sd[k].shape, generator=g))
        elif k.endswith("running_var"):
            sd[k].copy_(0.5 + torch.rand(sd[k].shape, generator=g))
        elif ".1.weight" in k or k == "initial_batchnorm.weight":
            sd[k].copy_(0.75 + 0.5 * torch.rand(sd[k].shape, generator=g))
        elif ".1.bias" in k or k == "initial_batchnorm.bias":
            sd[k].copy_(0.1 * torch.randn(sd[k].shape, generator=g))


def main():
    mods = rh.load("drs")
    with rh.argv_variant("drs"):
        rn = importlib.import_module("r_network")
        cfg = mods["config"].Config()
        net = rn.R_NETWORK(cfg, dict(mods["config"].hparams), 0).eval()
    digest = SW.state_dict_digest(net.state_dict())
    keys = [(k, tuple(v.shape), str(v.dtype)) for k, v in net.state_dict().items()]
    randomise_bn(net.state_dict(), 7)
    B, T = 2, 64
    _, _, noisy = O.synthetic_audio(B, 32 * (T - 1))
    spec = rh.reference_stft(noisy)
    noisy_mag = torch.abs(spec)
    noisy_phase = torch.atan2(spec.imag, spec.real + 10e-7)
    taps = {}
    hooks = [net.encoder[i].register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"enc{i}", o.detach().clone())) for i in (0, 3, 6)]
    with torch.no_grad():
        mask = net(noisy_mag)
    for h in hooks:
        h.remove()
    mp2w = lambda mag, ph: torch.istft(torch.nn.functional.pad(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), (0, 0, 0, 1)),  # noqa: E731
                                       n_fft=cfg.fft_size, hop_length=cfg.hop_length, win_length=cfg.window_length,
                                       window=cfg.window, normalized=cfg.normalise_stft)      # network_functions.py:140-150
    drs_noise = noisy_mag * mask                                                               # network_functions.py:301-302
    drs_clean = noisy_mag - drs_noise
    dr_clean = noisy_mag * mask                                                                # network_functions.py:341
    fp = lambda t: dict(mean_abs=float(t.abs().mean()), max_abs=float(t.abs().max()), head=t.reshape(-1)[:32].clone(), shape=tuple(t.shape))  # noqa: E731
    torch.save(dict(B=B, T=T, keys=keys, digest_seed0=digest, bn_seed=7, noisy_audio=noisy, mask=mask,
                    taps={k: fp(v) for k, v in taps.items()}, drs_clean_mag=drs_clean, dr_clean_mag=dr_clean,
                    drs_clean_audio=mp2w(drs_clean, noisy_phase), dr_clean_audio=mp2w(dr_clean, noisy_phase)),
               os.path.join(OUT, "rnet_drs_randbn_B2_T64.pt"))
    print("wrote rnet_drs_randbn_B2_T64.pt", len(keys), "keys", digest[:16])


if __name__ == "__main__":
    main()
