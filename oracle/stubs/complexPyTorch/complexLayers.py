"""complexPyTorch 0.3 `complexLayers` restated (see package docstring)."""
import torch
from torch import nn

from complexPyTorch.complexFunctions import apply_complex, complex_relu, complex_avg_pool2d

_SQRT2 = 1.4142135623730951


class ComplexReLU(nn.Module):
    def forward(self, x):
        return complex_relu(x)


class ComplexAvgPool2d(nn.Module):  # must exist: c_network.py:6 deletes the name
    def __init__(self, kernel_size, stride=None, padding=0, ceil_mode=False,
                 count_include_pad=True, divisor_override=None):
        super().__init__()
        self.args = (kernel_size, stride, padding, ceil_mode, count_include_pad, divisor_override)

    def forward(self, x):
        return complex_avg_pool2d(x, *self.args)


class ComplexConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0,
                 dilation=1, groups=1, bias=True):
        super().__init__()
        self.conv_r = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.conv_i = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)

    def forward(self, x):
        return apply_complex(self.conv_r, self.conv_i, x)


class ComplexConvTranspose2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 output_padding=0, groups=1, bias=True, dilation=1, padding_mode="zeros"):
        super().__init__()
        self.conv_tran_r = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding,
                                              output_padding, groups, bias, dilation, padding_mode)
        self.conv_tran_i = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding,
                                              output_padding, groups, bias, dilation, padding_mode)

    def forward(self, x):
        return apply_complex(self.conv_tran_r, self.conv_tran_i, x)


class ComplexLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.fc_r = nn.Linear(in_features, out_features)
        self.fc_i = nn.Linear(in_features, out_features)

    def forward(self, x):
        return apply_complex(self.fc_r, self.fc_i, x)


class _ComplexBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.track_running_stats = track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.empty(num_features, 3))
            self.bias = nn.Parameter(torch.empty(num_features, 2))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        if track_running_stats:
            self.register_buffer("running_mean", torch.zeros(num_features, dtype=torch.complex64))
            self.register_buffer("running_covar", torch.zeros(num_features, 3))
            self.running_covar[:, 0] = _SQRT2
            self.running_covar[:, 1] = _SQRT2
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.register_parameter("running_mean", None)
            self.register_parameter("running_covar", None)
            self.register_parameter("num_batches_tracked", None)
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
    def forward(self, x):
        eaf = 0.0
        if self.training and self.track_running_stats:
            if self.num_batches_tracked is not None:
                self.num_batches_tracked += 1
                eaf = 1.0 / float(self.num_batches_tracked) if self.momentum is None else self.momentum

        use_batch = self.training or (not self.training and not self.track_running_stats)
        if use_batch:
            mean = x.real.mean([0, 2, 3]).type(torch.complex64) + 1j * x.imag.mean([0, 2, 3]).type(torch.complex64)
        else:
            mean = self.running_mean
        if self.training and self.track_running_stats:
            with torch.no_grad():
                self.running_mean = eaf * mean + (1 - eaf) * self.running_mean

        x = x - mean[None, :, None, None]

        if use_batch:
            n = x.numel() / x.size(1)
            Crr = 1.0 / n * x.real.pow(2).sum(dim=[0, 2, 3]) + self.eps
            Cii = 1.0 / n * x.imag.pow(2).sum(dim=[0, 2, 3]) + self.eps
            Cri = (x.real.mul(x.imag)).mean(dim=[0, 2, 3])
        else:
            Crr = self.running_covar[:, 0] + self.eps
            Cii = self.running_covar[:, 1] + self.eps
            Cri = self.running_covar[:, 2]
        if self.training and self.track_running_stats:
            with torch.no_grad():
                self.running_covar[:, 0] = eaf * Crr * n / (n - 1) + (1 - eaf) * self.running_covar[:, 0]
                self.running_covar[:, 1] = eaf * Cii * n / (n - 1) + (1 - eaf) * self.running_covar[:, 1]
                self.running_covar[:, 2] = eaf * Cri * n / (n - 1) + (1 - eaf) * self.running_covar[:, 2]

        # inverse matrix square root of [[Crr, Cri], [Cri, Cii]]
        det = Crr * Cii - Cri.pow(2)
        s = torch.sqrt(det)
        t = torch.sqrt(Cii + Crr + 2 * s)
        ist = 1.0 / (s * t)
        Rrr = (Cii + s) * ist
        Rii = (Crr + s) * ist
        Rri = -Cri * ist
        b = lambda v: v[None, :, None, None]
        x = (b(Rrr) * x.real + b(Rri) * x.imag).type(torch.complex64) \
            + 1j * (b(Rii) * x.imag + b(Rri) * x.real).type(torch.complex64)
        if self.affine:
            w, c = self.weight, self.bias
            x = (b(w[:, 0]) * x.real + b(w[:, 2]) * x.imag + b(c[:, 0])).type(torch.complex64) \
                + 1j * (b(w[:, 2]) * x.real + b(w[:, 1]) * x.imag + b(c[:, 1])).type(torch.complex64)
        return x
