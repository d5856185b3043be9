"""dcsnet_b200 — B200-native (sm_100a) implementation of the DCS-Net forward hot path
(STFT -> complex encoder/decoder -> bounded mask / subtraction -> iSTFT) behind the reference's module API.

Host code is Python/PyTorch (device memory, streams, torch.distributed plumbing); all arithmetic on the path runs in
hand-written CUDA kernels reached through the C ABI in include/dcsnet.h.  There is no CPU or PyTorch-op fallback.
"""
from . import _lib, ops, packing, engine  # noqa: F401
from .engine import ForwardPlan, PackedNet  # noqa: F401

__all__ = ["ops", "packing", "engine", "ForwardPlan", "PackedNet"]
