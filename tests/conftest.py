import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(a, b):
    """max|a-b| / max|b| — the 'max relative error' of the north_star: max-abs error normalised by the
    max-abs reference (element-wise relative error is ill-defined at spectral nulls)."""
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.is_complex():
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_product_net(state="default"):
    """The product's C_NETWORK(config, hparams, seed=0) (+ seeded BN randomisation), eval mode, on CPU."""
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, config as cfg
    from oracle import synthetic_weights as SW
    net = c_network.C_NETWORK(cfg.config, cfg.hparams, 0)
    if state == "randbn":
        SW.randomise_bn_state(net.state_dict(), 7)
    return net.eval()
