"""CPU: the epoch-end glue of the Lightning layer (network_functions.py:450-498 epoch_end, c_network.py:304-335
validation_epoch_end, 374-398 test_epoch_end) against the REFERENCE's own functions executed through oracle/reference_harness.py
on the same step outputs and the same numpy seed: identical logger calls (tags, sample choice, order) and identical averaged metrics.
Skipped where /root/reference is absent (the GPU box)."""
import types

import numpy as np
import pytest
import torch

from oracle import reference_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


class _Recorder:
    def __init__(self):
        self.calls = []
        self.experiment = self

    def add_audio(self, tag, sample, step, sample_rate=None):
        self.calls.append(("audio", tag, np.asarray(sample).copy(), int(step), sample_rate))

    def add_scalar(self, tag, value, global_step=None):
        self.calls.append(("scalar", tag, float(value), global_step))


def _outputs(variant, n_batches=3, batch=4, length=50, prefix="val", seed=0):
    g = np.random.default_rng(seed)
    keys = ["clean", "predict_clean", "noise"] + (["predict_noise"] if variant in ("dcs", "drs") else []) + ["noisy"]
    outs = []
    for b in range(n_batches):
        audio = {k: g.standard_normal((batch, length)).astype(np.float32) for k in keys}
        m = {f"{prefix}_speech_loss": torch.tensor(float(g.standard_normal())), f"{prefix}_pesq": torch.tensor(float(g.uniform(1, 4))),
             f"{prefix}_stoi": torch.tensor(float(g.uniform(0, 1)))}
        if variant in ("dcs", "drs"):
            m[f"{prefix}_loss"], m[f"{prefix}_noise_loss"] = torch.tensor(float(g.standard_normal())), torch.tensor(float(g.standard_normal()))
        outs.append((audio, m))
    return outs


@pytest.mark.parametrize("variant", ["dcs", "dc"])
@pytest.mark.parametrize("hook", ["validation_epoch_end", "test_epoch_end"])
def test_epoch_end_hooks_equal_the_reference(variant, hook):
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network as ours_mod, config as ours_cfg
    mods = rh.load("dcs")
    prefix = "val" if hook.startswith("val") else "test"
    outs = _outputs(variant, prefix=prefix)
    results = []
    for which in ("reference", "ours"):
        rec = _Recorder()
        logged = []
        if which == "reference":
            with rh.argv_variant(variant):
                net = mods["c_network"].C_NETWORK(mods["config"].config, dict(mods["config"].hparams), 0)
        else:
            net = ours_mod.C_NETWORK(ours_cfg.config, dict(ours_cfg.hparams), 0)
            net.variant = variant
        net.config.val_log_sample_size = 2
        # what Lightning's Trainer would provide
        for name, value in (("logger", rec), ("current_epoch", 3), ("global_step", 17)):
            try:
                setattr(net, name, value)
            except AttributeError:
                setattr(type(net), name, value)
        net.log_dict = types.MethodType(lambda self, m, **k: logged.append(dict(m)), net)
        np.random.seed(5)
        with rh.argv_variant(variant):
            metrics = getattr(net, hook)(outs)
        results.append((rec.calls, metrics, logged))
    (c_ref, m_ref, l_ref), (c_our, m_our, l_our) = results
    assert len(c_ref) == len(c_our) > 0
    for a, b in zip(c_ref, c_our):
        assert a[0] == b[0] and a[1] == b[1] and a[3:] == b[3:], (a[1], b[1])
        assert np.array_equal(a[2], b[2])
    assert set(m_ref) == set(m_our)
    for k in m_ref:
        assert abs(float(m_ref[k]) - float(m_our[k])) <= 1e-6, k
    assert len(l_ref) == len(l_our) == 1 and set(l_ref[0]) == set(l_our[0])
