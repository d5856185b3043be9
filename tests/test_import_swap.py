"""Build container only: INTEGRATION.md's import swap, executed.  The reference's OWN `c_network.py` is imported unmodified
with `complexPyTorch` aliased to this package's drop-in modules (`dcsnet_b200.complexLayers` / `complexFunctions`), its
`C_NETWORK(config, hparams, seed)` constructor is run, and the resulting state_dict must be the reference's: same keys, same
order, same shapes / dtypes and — because module registration order and RNG consumption are the same — the same values as
the product's own `C_NETWORK` and as the reference built on the oracle's complexPyTorch restatement.

Runs in a subprocess (the alias must be in sys.modules before the reference is imported, and other tests import the
reference behind the oracle's stubs).  Skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys

import pytest

from oracle import reference_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")

SCRIPT = r"""
import sys, types, hashlib
sys.path.insert(0, {root!r})
import torch
import dcsnet_b200
from dcsnet_b200 import complexLayers, complexFunctions
pkg = types.ModuleType("complexPyTorch")
pkg.complexLayers, pkg.complexFunctions = complexLayers, complexFunctions
sys.modules["complexPyTorch"] = pkg
sys.modules["complexPyTorch.complexLayers"] = complexLayers
sys.modules["complexPyTorch.complexFunctions"] = complexFunctions
import torchaudio
if not hasattr(torchaudio, "set_audio_backend"):
    torchaudio.set_audio_backend = lambda *a, **k: None
sys.path.insert(0, {stubs!r})          # pytorch_lightning / pypesq / pystoi import stubs (complexPyTorch is already aliased)
sys.path.insert(0, {ref!r})
sys.argv = ["swap", "dcs", "0"]
import c_network as ref_cn, config as ref_cfg
assert ref_cn.ComplexConv2d is complexLayers.ComplexConv2d, "the reference did not pick up the drop-in layers"
net = ref_cn.C_NETWORK(ref_cfg.config, ref_cfg.hparams, 0)
from dcsnet_b200 import c_network as my_cn, config as my_cfg
mine = my_cn.C_NETWORK(my_cfg.config, my_cfg.hparams, 0)
a, b = net.state_dict(), mine.state_dict()
assert list(a.keys()) == list(b.keys()), "state_dict keys / order differ"
for k in a:
    assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
    assert torch.equal(a[k], b[k]), k
print("SWAP_OK", len(a), sum(v.numel() for v in net.parameters()))
"""


def test_reference_c_network_builds_on_the_dropin_layers():
    code = SCRIPT.format(root=ROOT, stubs=os.path.join(ROOT, "oracle", "stubs"), ref=rh.REFERENCE_ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SWAP_OK 246 2912707" in r.stdout, r.stdout
