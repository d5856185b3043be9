"""Real network path (r_network.py): the CPU oracle against vectors produced by the reference itself,
and the drop-in R_NETWORK parameter container (state_dict keys / shapes / seed-0 weights identical to the reference)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rnet_oracle as RO, dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402
from oracle.make_golden_rnet import randomise_bn  # noqa: E402
import dcsnet_b200 as D  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "rnet_drs_randbn_B2_T64.pt")


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def product_net():
    from dcsnet_b200 import r_network, config as C
    return r_network.R_NETWORK(C.Config(), dict(C.hparams), 0).eval()


def test_rnetwork_state_dict_and_seed0_weights_identical_to_reference():
    g = torch.load(GOLDEN)
    net = product_net()
    sd = net.state_dict()
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()] == g["keys"]
    assert SW.state_dict_digest(sd) == g["digest_seed0"]            # bit-identical random-init weights


def test_rnet_oracle_matches_reference_golden():
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    sd = net.state_dict()
    spec = O.stft(g["noisy_audio"])
    taps = {}
    mask = RO.r_network_forward(sd, torch.abs(spec), taps=taps)
    assert rel_err(mask, g["mask"]) <= 2e-6
    for k, fp in g["taps"].items():
        assert tuple(taps[k].shape) == fp["shape"]
        assert abs(float(taps[k].abs().max()) - fp["max_abs"]) <= 1e-5 * fp["max_abs"]
        assert rel_err(taps[k].reshape(-1)[:32], fp["head"]) <= 1e-5
    for variant in ("drs", "dr"):
        out = RO.enhance_spec(sd, spec, variant)
        assert rel_err(out["clean_mag"], g[f"{variant}_clean_mag"]) <= 2e-6
        assert rel_err(out["clean_audio"], g[f"{variant}_clean_audio"]) <= 2e-6


def test_channel_attention_uses_only_the_max_pool_branch():
    """r_network.py:23-24: `out = avg_out_fc + max_out_fc` is overwritten by `out = max_out_fc`."""
    g = torch.Generator().manual_seed(0)
    sd = {"a.fc.0.weight": torch.randn(2, 16, 1, 1, generator=g), "a.fc.2.weight": torch.randn(16, 2, 1, 1, generator=g)}
    x = torch.randn(3, 16, 5, 7, generator=g)
    got = RO.channel_attention(x, sd, "a.")
    y = x.clone()
    flat = y.view(3, 16, -1)
    flat.scatter_(-1, flat.argmin(-1, keepdim=True), -100.0)        # changes every average, no maximum
    assert torch.equal(got, RO.channel_attention(y, sd, "a."))


def test_product_forward_fails_loudly_without_fallback():
    net = product_net()
    with pytest.raises(NotImplementedError, match="no CPU"):
        net(torch.rand(2, 256, 64))
