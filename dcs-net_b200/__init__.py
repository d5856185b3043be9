"""dcsnet_b200 — B200-native (sm_100a) implementation of the DCS-Net forward hot path
(STFT -> complex encoder/decoder -> bounded mask / subtraction -> iSTFT) behind the reference's module API.

Host code is Python/PyTorch (device memory, streams, torch.distributed plumbing); all arithmetic on the path runs in
hand-written CUDA kernels reached through the C ABI in include/dcsnet.h.  There is no CPU or PyTorch-op fallback.

Reference-facing modules (same names as the reference's files):
    dcsnet_b200.complexLayers / complexFunctions   <- complexPyTorch 0.3 (c_network.py:5-7)
    dcsnet_b200.network_functions                  <- network_functions.py (layer / mask / iSTFT part)
    dcsnet_b200.c_network                          <- c_network.py (ComplexLSTM, attention, C_NETWORK)
    dcsnet_b200.config                             <- config.py (hparams, Config)
"""
from . import _lib, ops, packing, engine, rengine, pipeline, frontend  # noqa: F401
from .frontend import GpuFrontEnd  # noqa: F401
from .engine import ForwardPlan, PackedNet  # noqa: F401
from .rengine import RealForwardPlan, PackedRealNet  # noqa: F401
from .pipeline import RealEnhancer  # noqa: F401
from .pipeline import Enhancer, shard_range, split_windows, frames_for, WINDOW_4S, WINDOW_REF  # noqa: F401
from . import complexFunctions, complexLayers, network_functions, config, c_network  # noqa: F401
from .c_network import C_NETWORK  # noqa: F401

__all__ = ["ops", "packing", "engine", "pipeline", "frontend", "GpuFrontEnd", "ForwardPlan", "PackedNet", "Enhancer", "C_NETWORK", "RealForwardPlan", "PackedRealNet", "RealEnhancer",
           "complexLayers", "complexFunctions", "network_functions", "c_network", "config"]
