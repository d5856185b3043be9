"""Run the reference's own files UNMODIFIED as the oracle (build container only).

TEST INFRASTRUCTURE.  `/root/reference` is read-only and exists only in the build container;
nothing that runs on the GPU box may import this module (see `available()`).

What is stubbed (oracle/stubs/): pytorch_lightning, pypesq, pystoi, complexPyTorch (restated,
SURVEY Appendix A) and the removed `torchaudio.set_audio_backend` (config.py:10).  `sys.argv[1]`
is read inside the reference's library code (network_functions.py:170...), so it is set here.
"""
import importlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("DCSNET_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_cache = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "c_network.py"))


def load(variant="dcs"):
    """Import config / network_functions / c_network from the reference tree. Returns a namespace dict."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    if "mods" not in _cache:
        import torchaudio
        if not hasattr(torchaudio, "set_audio_backend"):
            torchaudio.set_audio_backend = lambda *a, **k: None
        for p in (_STUBS, REFERENCE_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        argv = sys.argv
        sys.argv = ["oracle", variant, "0"]
        try:
            mods = {n: importlib.import_module(n) for n in ("network_functions", "config", "c_network")}
        finally:
            sys.argv = argv
        _cache["mods"] = mods
    return _cache["mods"]


class argv_variant:
    """Context manager: the reference reads sys.argv[1] ∈ {dcs,drs,dc,dr} inside library functions."""

    def __init__(self, variant):
        self.variant = variant

    def __enter__(self):
        self._old = sys.argv
        sys.argv = ["oracle", self.variant, "0"]

    def __exit__(self, *a):
        sys.argv = self._old


def build_c_network(seed=0, randomise_bn=True, hparam_overrides=None):
    """`C_NETWORK(config, hparams, seed).eval()` exactly as the reference builds it (c_network.py:88-171)."""
    m = load()
    hp = dict(m["config"].hparams)
    if hparam_overrides:
        hp.update(hparam_overrides)
    net = m["c_network"].C_NETWORK(m["config"].config, hp, seed)
    if randomise_bn:
        randomise_bn_state(net)
    return net.eval()


def randomise_bn_state(net, seed=7):
    """Seeded non-trivial BN buffers/affine (default BN state is isotropic and hides bugs)."""
    g = torch.Generator().manual_seed(seed)
    sd = net.state_dict()
    with torch.no_grad():
        for k in sd:
            if k.endswith("running_covar"):
                c = sd[k].shape[0]
                sd[k][:, 0] = 0.5 + torch.rand(c, generator=g)
                sd[k][:, 1] = 0.5 + torch.rand(c, generator=g)
                sd[k][:, 2] = 0.6 * torch.rand(c, generator=g) - 0.3
                base = k[: -len("running_covar")]
                sd[base + "running_mean"].copy_(torch.complex(0.2 * torch.randn(c, generator=g),
                                                              0.2 * torch.randn(c, generator=g)))
                w = sd[base + "weight"]
                w[:, 0] = 1.0 + 0.5 * torch.rand(c, generator=g)
                w[:, 1] = 1.0 + 0.5 * torch.rand(c, generator=g)
                w[:, 2] = 0.4 * torch.rand(c, generator=g) - 0.2
                sd[base + "bias"].copy_(0.1 * torch.randn(c, 2, generator=g))
    return net


def mag_phase_2_wave_cpu(mag, phase, config):
    """Device-agnostic twin of network_functions.py:140-150 (the original hard-codes a cuda window, line 147)."""
    comp = torch.complex(mag * torch.cos(phase), mag * torch.sin(phase))
    comp = torch.nn.functional.pad(comp, (0, 0, 0, 1))
    return torch.istft(comp, n_fft=config.fft_size, hop_length=config.hop_length, win_length=config.window_length,
                       window=config.window.to(comp.device), normalized=config.normalise_stft)


def reference_stft(audio):
    """data.py:112-134 applied to a batch (B, L) -> (B, 256, T) complex64."""
    cfg = load()["config"].config
    spec = torch.stft(audio, n_fft=cfg.fft_size, hop_length=cfg.hop_length, win_length=cfg.window_length,
                      window=cfg.window, return_complex=True, normalized=cfg.normalise_stft)
    return spec[..., 1:int(cfg.fft_size / 2) + 1, :]


def reference_enhance(net, noisy_spec, variant="dcs"):
    """The combine lines of network_functions.py:388-401 (dcs) / 428-436 (dc), executed with the reference's
    own functions.  Returns dict(mask, noise_spec, clean_spec, noise_audio, clean_audio)."""
    m = load()
    nf, cfg = m["network_functions"], m["config"].config
    hp = net.hparams
    eps = hp["atan2_eps"]
    with torch.no_grad(), argv_variant(variant):
        out = net(noisy_spec)
        mask = nf.bound_cRM(out, hp)
        prod = nf.complex_mat_mult(noisy_spec, mask)
        if variant == "dcs":
            noise_spec, clean_spec = prod, noisy_spec - prod
        else:
            noise_spec, clean_spec = None, prod
        w = lambda s: mag_phase_2_wave_cpu(torch.abs(s), torch.atan2(s.imag, s.real + eps), cfg)
        return dict(net_out=out, mask=mask, noise_spec=noise_spec, clean_spec=clean_spec,
                    noise_audio=None if noise_spec is None else w(noise_spec), clean_audio=w(clean_spec))
