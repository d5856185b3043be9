"""Import stub (TEST INFRASTRUCTURE ONLY) standing in for pytorch_lightning==1.5.6.

The reference (`/root/reference/c_network.py:2-4`, `network_functions.py:4`, `data.py:7`)
imports pytorch_lightning only for `LightningModule`, `seed_everything`, `Callback` and
the Trainer glue.  None of that is arithmetic on the hot path, so a minimal stand-in is
enough to let the reference files import UNMODIFIED in the build container and act as the
parity oracle.  Nothing under `dcs-net_b200/` may import this.
"""
import random as _random

import numpy as _np
import torch as _torch


class _HParams(dict):
    """dict with attribute access, like Lightning's AttributeDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(_torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        object.__setattr__(self, "_hp", _HParams())
        self.current_epoch = 0
        self.trainer = None
        self.logger = None

    @property
    def hparams(self):
        return self._hp

    def save_hyperparameters(self, *a, **k):
        return None

    def log_dict(self, *a, **k):
        return None

    def log(self, *a, **k):
        return None


class Callback:
    pass


class Trainer:  # never used by the oracle; present so `from pytorch_lightning import Trainer` resolves
    def __init__(self, *a, **k):
        raise RuntimeError("pytorch_lightning stub: Trainer is out of scope for the oracle")


class _Loggers:
    class TensorBoardLogger:
        def __init__(self, *a, **k):
            pass


loggers = _Loggers()


def seed_everything(seed=None, workers=False):
    """Same RNG side effects as pl.seed_everything (python, numpy, torch, cuda)."""
    seed = int(seed)
    _random.seed(seed)
    _np.random.seed(seed)
    _torch.manual_seed(seed)
    if _torch.cuda.is_available():
        _torch.cuda.manual_seed_all(seed)
    return seed
