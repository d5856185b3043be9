"""Build container only: the oracle restatement against the reference's own files executed unmodified (behind the
import stubs of oracle/stubs).  Skipped where /root/reference does not exist (the GPU box)."""
import pytest
import torch

from oracle import reference_harness as rh, dcsnet_oracle as O, synthetic_weights as SW

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


def test_port_is_bit_exact_vs_reference():
    net = rh.build_c_network(0, randomise_bn=True)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    _, _, noisy = O.synthetic_audio(2, 32 * 63)
    spec = rh.reference_stft(noisy)
    assert torch.equal(O.stft(noisy), spec)
    for variant in ("dcs", "dc"):
        ref = rh.reference_enhance(net, spec, variant)
        got = O.enhance_spec(sd, spec, variant)
        for k in ("net_out", "mask", "clean_spec"):
            assert torch.equal(got[k], ref[k]), (variant, k)
        assert torch.equal(O.spec_to_wave(got["clean_spec"]), ref["clean_audio"])


def test_product_constructor_reproduces_reference_state_dict():
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, config as cfg
    mine = c_network.C_NETWORK(cfg.config, cfg.hparams, 0).state_dict()
    ref = rh.build_c_network(0, randomise_bn=False).state_dict()
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert mine[k].dtype == ref[k].dtype and torch.equal(mine[k], ref[k]), k


def test_bn_randomiser_matches_harness():
    a = rh.build_c_network(0, randomise_bn=True).state_dict()
    b = rh.build_c_network(0, randomise_bn=False).state_dict()
    SW.randomise_bn_state(b, 7)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_reference_batch_independence():
    """The only 'test' idea in the reference (CheckBatchGradient, network_functions.py:517-532): samples are independent."""
    net = rh.build_c_network(0)
    _, _, noisy = O.synthetic_audio(3, 32 * 31)
    spec = rh.reference_stft(noisy)
    with torch.no_grad():
        full, part = net(spec), net(spec[:2])
    assert torch.equal(full[:2], part)
