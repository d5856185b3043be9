"""Drop-in replacements for the `complexPyTorch.complexLayers` classes the reference imports
(/root/reference/c_network.py:5, config.py:5; complexPyTorch==0.3, requirements.txt:38).

Same class names, constructor signatures, sub-module attribute names and parameter shapes — hence the same
state_dict keys and, because the same torch initialisers run in the same order, the same
random-init weights for a given seed.  `forward` runs hand-written sm_100a kernels through the C ABI
(include/dcsnet.h) on CUDA tensors; there is no CPU / ATen fallback: CPU tensors raise.

Inference (eval-mode) semantics only in this round; train-mode batch statistics / autograd are SURVEY §8f rank 2.
"""
import torch
from torch import nn

from . import _lib as L
from . import ops, packing
from .complexFunctions import to_cl, from_cl, complex_relu  # noqa: F401

_SQRT2 = 1.4142135623730951


def _version_key(*params):
    return tuple((p.data_ptr(), p._version) for p in params if p is not None)


class _PackedCache:
    """Re-pack GEMM operands only when a parameter tensor changed (in-place update or re-assignment)."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, key, make):
        if key != self.key:
            self.key, self.val = key, make()
        return self.val


class ComplexReLU(nn.Module):
    def forward(self, input):
        return complex_relu(input)


class ComplexAvgPool2d(nn.Module):
    """Exists only because c_network.py:6 deletes the name after the star-import; never instantiated by the path."""

    def __init__(self, kernel_size, stride=None, padding=0, ceil_mode=False, count_include_pad=True,
                 divisor_override=None):
        super().__init__()

    def forward(self, input):
        raise NotImplementedError("ComplexAvgPool2d is not on the DCS-Net hot path (deleted at c_network.py:6)")


class ComplexConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        self.conv_r = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.conv_i = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self._cache = _PackedCache()

    def _packed(self, device):
        r, i = self.conv_r, self.conv_i
        k = r.kernel_size
        if r.dilation != (1, 1) or r.groups != 1 or r.padding != (k[0] // 2, k[1] // 2):
            raise NotImplementedError("dcsnet_b200.ComplexConv2d: only dilation=1, groups=1, padding=k//2 "
                                      "(every configuration in config.py:83-99 and the attention blocks)")
        key = _version_key(r.weight, i.weight, r.bias, i.bias) + (str(device),)
        return self._cache.get(key, lambda: packing.PackedConv(r.weight, i.weight, r.bias, i.bias, stride=r.stride,
                                                               device=device))

    def forward(self, input):
        L.require_cuda(input)
        x = to_cl(input)
        pk = self._packed(x.device)
        oh, ow = ops.conv_out_hw(pk, x.shape[1], x.shape[2])
        y = torch.empty(x.shape[0], oh, ow, pk.cout, 2, dtype=torch.float32, device=x.device)
        return from_cl(ops.cconv(pk, x, None, y))


class ComplexConvTranspose2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1,
                 bias=True, dilation=1, padding_mode='zeros'):
        super().__init__()
        self.conv_tran_r = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding, output_padding,
                                              groups, bias, dilation, padding_mode)
        self.conv_tran_i = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding, output_padding,
                                              groups, bias, dilation, padding_mode)
        self._cache = _PackedCache()

    def _packed(self, device):
        r, i = self.conv_tran_r, self.conv_tran_i
        k = r.kernel_size
        if (r.stride != (1, 1) or r.dilation != (1, 1) or r.groups != 1 or r.output_padding != (0, 0)
                or r.padding != (k[0] // 2, k[1] // 2) or k[0] % 2 == 0 or k[1] % 2 == 0):
            raise NotImplementedError("dcsnet_b200.ComplexConvTranspose2d: only odd k, stride 1, padding k//2 "
                                      "(config.py:84,92-100), i.e. the flipped same-size convolution")
        key = _version_key(r.weight, i.weight, r.bias, i.bias) + (str(device),)
        return self._cache.get(key, lambda: packing.PackedConv(r.weight, i.weight, r.bias, i.bias, transposed=True,
                                                               device=device))

    def forward(self, input):
        L.require_cuda(input)
        x = to_cl(input)
        pk = self._packed(x.device)
        y = torch.empty(x.shape[0], x.shape[1], x.shape[2], pk.cout, 2, dtype=torch.float32, device=x.device)
        return from_cl(ops.cconv(pk, x, None, y))


class ComplexLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.fc_r = nn.Linear(in_features, out_features)
        self.fc_i = nn.Linear(in_features, out_features)
        self._cache = _PackedCache()

    def forward(self, input):
        L.require_cuda(input)
        r, i = self.fc_r, self.fc_i
        key = _version_key(r.weight, i.weight, r.bias, i.bias) + (str(input.device),)
        pk = self._cache.get(key, lambda: packing.PackedConv(r.weight[:, :, None, None], i.weight[:, :, None, None],
                                                             r.bias, i.bias, device=input.device))
        lead = input.shape[:-1]
        x = torch.view_as_real(input.contiguous()).reshape(1, 1, -1, input.shape[-1], 2)
        y = torch.empty(1, 1, x.shape[2], pk.cout, 2, dtype=torch.float32, device=input.device)
        ops.cconv(pk, x, None, y)
        return torch.view_as_complex(y).reshape(*lead, pk.cout)


class _ComplexBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.track_running_stats = track_running_stats
        if self.affine:
            self.weight = nn.Parameter(torch.empty(num_features, 3))
            self.bias = nn.Parameter(torch.empty(num_features, 2))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)
        if self.track_running_stats:
            self.register_buffer('running_mean', torch.zeros(num_features, dtype=torch.complex64))
            self.register_buffer('running_covar', torch.zeros(num_features, 3))
            self.running_covar[:, 0] = _SQRT2
            self.running_covar[:, 1] = _SQRT2
            self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        else:
            self.register_parameter('running_mean', None)
            self.register_parameter('running_covar', None)
            self.register_parameter('num_batches_tracked', None)
        self.reset_parameters()

    def reset_running_stats(self):
        if self.track_running_stats:
            self.running_mean.zero_()
            self.running_covar.zero_()
            self.running_covar[:, 0] = _SQRT2
            self.running_covar[:, 1] = _SQRT2
            self.num_batches_tracked.zero_()

    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:
            with torch.no_grad():
                self.weight[:, :2].fill_(_SQRT2)
                self.weight[:, 2].zero_()
                self.bias.zero_()


class ComplexBatchNorm2d(_ComplexBatchNorm):
    """Eval mode = per-channel 2x2 affine, applied by dcs_cbn_apply."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._cache = _PackedCache()

    def folded_affine(self, device):
        if not (self.affine and self.track_running_stats):
            raise NotImplementedError("dcsnet_b200.ComplexBatchNorm2d: affine=True, track_running_stats=True only")
        key = _version_key(self.weight, self.bias, self.running_mean, self.running_covar) + (str(device),)
        return self._cache.get(key, lambda: packing.affine6(*packing.bn_affine(
            self.weight.detach().cpu(), self.bias.detach().cpu(), self.running_mean.cpu(), self.running_covar.cpu(),
            self.eps)).to(device))

    def forward(self, input):
        if self.training:
            raise NotImplementedError("dcsnet_b200.ComplexBatchNorm2d: the layer-wise module has no autograd twin; train-mode batch statistics "
                                      "(forward and backward) are dcsnet_b200.train_ops.cbn_train_fwd / cbn_train_bwd, driven by "
                                      "train_engine.TrainStep; call .eval() for the folded inference form")
        L.require_cuda(input)
        x = to_cl(input)
        return from_cl(ops.cbn_apply(x, self.folded_affine(x.device)))
