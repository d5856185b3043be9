"""Generate tests/golden/*.pt by EXECUTING THE REFERENCE ITSELF (build container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden

The reference files under /root/reference are imported unmodified behind the stubs in oracle/stubs (see
oracle/reference_harness.py).  Weights are not stored (11.65 MB): they are `C_NETWORK(config, hparams, seed=0)`,
reproducible from the seed; the fixture stores their sha256 so a consumer can prove it rebuilt identical weights.
Two weight states are recorded: 'default' (the reference's init as is) and 'randbn' (same weights with seeded
non-trivial BN buffers/affine, oracle/synthetic_weights.randomise_bn_state(seed=7)).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_harness as rh, dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TAPS = ["enc0", "enc3", "enc6", "lstm", "fc", "skip0", "dec0", "dec3", "dec5", "dec6"]


def tap_summary(t):
    """Small fingerprint of an activation: mean |x|, max |x| and the first 32 values in memory order."""
    a = torch.view_as_real(t.detach().contiguous()).reshape(-1)
    return dict(mean_abs=float(a.abs().mean()), max_abs=float(a.abs().max()), head=a[:32].clone(), shape=tuple(t.shape))


def run_case(B, T, state, variant="dcs"):
    net = rh.build_c_network(0, randomise_bn=False)
    if state == "randbn":
        sd = net.state_dict()
        SW.randomise_bn_state(sd, 7)
    net.eval()
    digest = SW.state_dict_digest(net.state_dict())
    L = 32 * (T - 1)
    clean, noise, noisy = O.synthetic_audio(B, L)
    spec = rh.reference_stft(noisy)
    # per-layer taps via forward hooks on the reference modules
    taps = {}
    hooks = []
    for i in (0, 3, 6):
        hooks.append(net.encoder[i].register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"enc{i}", o)))
    hooks.append(net.lstm.register_forward_hook(lambda m, a, o: taps.__setitem__("lstm", o)))
    hooks.append(net.fc.register_forward_hook(lambda m, a, o: taps.__setitem__("fc", o)))
    r = rh.reference_enhance(net, spec, variant)
    for h in hooks:
        h.remove()
    out = dict(B=B, T=T, state=state, variant=variant, weights_sha256=digest, audio_seed=1234,
               noisy_audio=noisy, net_out=r["net_out"], clean_spec=r["clean_spec"], clean_audio=r["clean_audio"],
               taps={k: tap_summary(v) for k, v in taps.items()})
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = [(2, 64, "default", "dcs"), (2, 64, "randbn", "dcs"), (1, 32, "randbn", "dcs"), (2, 32, "randbn", "dc")]
    for B, T, state, variant in cases:
        g = run_case(B, T, state, variant)
        path = os.path.join(OUT, f"cnet_{variant}_{state}_B{B}_T{T}.pt")
        torch.save(g, path)
        print(path, os.path.getsize(path) // 1024, "KiB", g["weights_sha256"][:16])
    # STFT / iSTFT vectors straight from torch.stft / torch.istft with the reference's config (data.py:112-134,
    # network_functions.py:140-150), including a ragged T (T % 16 != 0) and the minimum length
    g = torch.Generator().manual_seed(99)
    vec = {}
    for name, (B, L) in dict(a=(2, 2016), b=(1, 32 * 40), c=(1, 512)).items():
        audio = 0.3 * torch.randn(B, L, generator=g)
        spec = rh.reference_stft(audio)
        s = torch.complex(torch.randn(spec.shape, generator=g), torch.randn(spec.shape, generator=g)) * 0.2
        cfg = rh.load()["config"].config
        eps = rh.load()["config"].hparams["atan2_eps"]
        wave = rh.mag_phase_2_wave_cpu(torch.abs(s), torch.atan2(s.imag, s.real + eps), cfg)
        vec[name] = dict(audio=audio, spec=spec, ispec=s, iwave=wave)
    path = os.path.join(OUT, "stft_istft.pt")
    torch.save(vec, path)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
