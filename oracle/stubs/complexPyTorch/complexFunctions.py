"""complexPyTorch 0.3 `complexFunctions` restated (see package docstring)."""
import torch
import torch.nn.functional as F


def apply_complex(fr, fi, x, dtype=torch.complex64):
    # (fr + j fi)(x.re + j x.im); every real module keeps its OWN bias, hence the
    # effective bias (b_r - b_i) + j (b_r + b_i)  (SURVEY Appendix A1 / D1).
    re = fr(x.real) - fi(x.imag)
    im = fr(x.imag) + fi(x.real)
    return re.type(dtype) + 1j * im.type(dtype)


def complex_relu(x):
    return F.relu(x.real).type(torch.complex64) + 1j * F.relu(x.imag).type(torch.complex64)


def complex_matmul(A, B):
    re = torch.matmul(A.real, B.real) - torch.matmul(A.imag, B.imag)
    im = torch.matmul(A.real, B.imag) + torch.matmul(A.imag, B.real)
    return re.type(torch.complex64) + 1j * im.type(torch.complex64)


def complex_upsample(input, size=None, scale_factor=None, mode="nearest",
                     align_corners=None, recompute_scale_factor=None):
    kw = dict(size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners,
              recompute_scale_factor=recompute_scale_factor)
    re = F.interpolate(input.real, **kw)
    im = F.interpolate(input.imag, **kw)
    return re.type(torch.complex64) + 1j * im.type(torch.complex64)


def complex_avg_pool2d(x, *a, **k):
    return F.avg_pool2d(x.real, *a, **k).type(torch.complex64) + 1j * F.avg_pool2d(x.imag, *a, **k).type(torch.complex64)
