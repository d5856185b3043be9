"""Drop-in replacements for the `complexPyTorch.complexFunctions` names the reference uses
(complex_upsample, complex_relu at /root/reference/c_network.py:7; complex_matmul at misc.py:20), plus the
layout helpers shared by the layer classes.  CUDA tensors only."""
import torch

from . import _lib as L
from . import ops


def to_cl(x):
    """complex64 NCHW (any strides) -> float32 (B,H,W,C,2) channels-last view (copy only if not already NHWC)."""
    if x.dtype != torch.complex64:
        raise RuntimeError(f"expected a complex64 tensor, got {x.dtype}")
    if x.dim() != 4:
        raise RuntimeError("expected a (B,C,H,W) tensor")
    return torch.view_as_real(x.permute(0, 2, 3, 1).contiguous())


def from_cl(y):
    """float32 (B,H,W,C,2) -> complex64 tensor of logical shape (B,C,H,W) in channels_last memory format."""
    return torch.view_as_complex(y).permute(0, 3, 1, 2)


_IDENT = {}


def _identity_affine(C, device):
    key = (C, str(device))
    if key not in _IDENT:
        _IDENT[key] = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0], device=device).repeat(C, 1).contiguous()
    return _IDENT[key]


def _eltwise_act(x, act):
    L.require_cuda(x)
    if x.dim() == 4:
        xr = to_cl(x)
        return from_cl(ops.cbn_apply(xr, _identity_affine(xr.shape[3], x.device), act=act))
    xr = torch.view_as_real(x.contiguous()).reshape(-1, 1, 2)
    y = ops.cbn_apply(xr, _identity_affine(1, x.device), act=act)
    return torch.view_as_complex(y.reshape(*x.shape, 2))


def complex_relu(input):
    return _eltwise_act(input, L.ACT_RELU)


def complex_upsample(input, size=None, scale_factor=None, mode='nearest', align_corners=None,
                     recompute_scale_factor=None):
    """Nearest-neighbour up-sampling by integer factors (the only mode the reference uses: config.py:105-106)."""
    if mode != 'nearest' or size is not None or scale_factor is None:
        raise NotImplementedError("dcsnet_b200.complex_upsample: mode='nearest' with an integer scale_factor only")
    sf = scale_factor if isinstance(scale_factor, (tuple, list)) else (scale_factor, scale_factor)
    if any(int(s) != s or s < 1 for s in sf):
        raise NotImplementedError("dcsnet_b200.complex_upsample: integer scale factors only")
    L.require_cuda(input)
    return from_cl(ops.upsample_nearest(to_cl(input), (int(sf[0]), int(sf[1]))))


def complex_matmul(A, B):
    raise NotImplementedError("complex_matmul is only referenced by the dead scratch file misc.py:20 "
                              "(SURVEY §2 row 14); it is not on the DCS-Net hot path")
