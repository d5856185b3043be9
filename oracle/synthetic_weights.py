"""Seeded random C_NETWORK state_dict in the reference's key/shape layout.  TEST INFRASTRUCTURE.

Independent of both the reference classes and the product's module classes, so kernel parity tests can run on the GPU
box (where /root/reference does not exist) without trusting either.  Scales follow the reference's initialisers
(xavier_uniform on conv/linear weights, torch defaults elsewhere); BN buffers/affine are randomised (SPD covariance)
because the default BN state is isotropic and hides bugs.
"""
import math

import torch

CHANNELS = [1, 16, 32, 64, 128, 256, 256, 256]
KERNEL_E = [7, 7, 5, 5, 3, 3, 3]


def _u(g, shape, bound):
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _xavier(g, shape):
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    return _u(g, shape, math.sqrt(6.0 / (fan_in + fan_out)))


def _bn(sd, p, c, g):
    sd[p + "weight"] = torch.stack([1.0 + 0.5 * torch.rand(c, generator=g), 1.0 + 0.5 * torch.rand(c, generator=g),
                                    0.4 * torch.rand(c, generator=g) - 0.2], 1)
    sd[p + "bias"] = 0.1 * torch.randn(c, 2, generator=g)
    sd[p + "running_mean"] = torch.complex(0.2 * torch.randn(c, generator=g), 0.2 * torch.randn(c, generator=g))
    sd[p + "running_covar"] = torch.stack([0.5 + torch.rand(c, generator=g), 0.5 + torch.rand(c, generator=g),
                                           0.6 * torch.rand(c, generator=g) - 0.3], 1)
    sd[p + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _cconv(sd, p, names, shape, g, bias=True):
    for n in names:
        sd[f"{p}{n}.weight"] = _xavier(g, shape)
        if bias:
            rf = shape[2] * shape[3] if len(shape) == 4 else 1
            fan_in = shape[1] * rf
            sd[f"{p}{n}.bias"] = _u(g, (shape[0] if "tran" not in n else shape[1],), 1.0 / math.sqrt(fan_in))


def _attention(sd, p_ca, p_sa, c, g, ratio=16):
    r = max(c // ratio, 1)
    _cconv(sd, p_ca + "fc.0.", ("conv_r", "conv_i"), (r, c, 1, 1), g, bias=False)
    _cconv(sd, p_ca + "fc.2.", ("conv_r", "conv_i"), (c, r, 1, 1), g, bias=False)
    _cconv(sd, p_sa + "conv1.", ("conv_r", "conv_i"), (1, 2, 7, 7), g, bias=False)


def make_state_dict(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    L = 7
    for i in range(L):
        cin = 1 if i == 0 else CHANNELS[i] // 2
        cout = CHANNELS[i + 1] // 2
        k = KERNEL_E[i]
        _cconv(sd, f"encoder.{i}.0.", ("conv_r", "conv_i"), (cout, cin, k, k), g)
        _bn(sd, f"encoder.{i}.1.", cout, g)
    for i in range(L):
        cin = CHANNELS[L - i]          # d + skip complex channels
        half = cin // 2
        cout = max(CHANNELS[L - 1 - i] // 2, 1)
        p = f"decoder.{i}." if i == L - 1 else f"decoder.{i}.0."
        _cconv(sd, p, ("conv_tran_r", "conv_tran_i"), (cin, cout, 3, 3), g)
        if i != L - 1:
            _bn(sd, f"decoder.{i}.1.", cout, g)
        _attention(sd, f"decoder_attention.{2 * i}.", f"decoder_attention.{2 * i + 1}.", cout, g)
        _attention(sd, f"skip_attention.{2 * i}.", f"skip_attention.{2 * i + 1}.", half, g)
    _bn(sd, "initial_batchnorm.", 1, g)
    h, d = 64, 128
    for n in ("real_lstm", "imag_lstm"):
        for l in range(2):
            for s in ("", "_reverse"):
                b = 1.0 / math.sqrt(h)
                sd[f"lstm.{n}.weight_ih_l{l}{s}"] = _u(g, (4 * h, d), b)
                sd[f"lstm.{n}.weight_hh_l{l}{s}"] = _u(g, (4 * h, h), b)
                sd[f"lstm.{n}.bias_ih_l{l}{s}"] = _u(g, (4 * h,), b)
                sd[f"lstm.{n}.bias_hh_l{l}{s}"] = _u(g, (4 * h,), b)
    for n in ("fc_r", "fc_i"):
        sd[f"fc.{n}.weight"] = _xavier(g, (128, 128))
        sd[f"fc.{n}.bias"] = _u(g, (128,), 1.0 / math.sqrt(128))
    return sd


def randomise_bn_state(sd, seed=7):
    """In-place seeded non-trivial BN buffers/affine on a reference-format state_dict: covariance kept
    SPD (Crr, Cii in [0.5, 1.5], |Cri| <= 0.3).  Iterates keys in state_dict order, so it is reproducible for both the
    reference's and the product's C_NETWORK (identical key order)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k in list(sd.keys()):
            if k.endswith("running_covar"):
                c = sd[k].shape[0]
                base = k[: -len("running_covar")]
                sd[k][:, 0] = 0.5 + torch.rand(c, generator=g)
                sd[k][:, 1] = 0.5 + torch.rand(c, generator=g)
                sd[k][:, 2] = 0.6 * torch.rand(c, generator=g) - 0.3
                sd[base + "running_mean"].copy_(torch.complex(0.2 * torch.randn(c, generator=g), 0.2 * torch.randn(c, generator=g)))
                w = sd[base + "weight"]
                w[:, 0] = 1.0 + 0.5 * torch.rand(c, generator=g)
                w[:, 1] = 1.0 + 0.5 * torch.rand(c, generator=g)
                w[:, 2] = 0.4 * torch.rand(c, generator=g) - 0.2
                sd[base + "bias"].copy_(0.1 * torch.randn(c, 2, generator=g))
    return sd


def state_dict_digest(sd):
    """sha256 over keys, shapes and raw bytes — pins 'identical random-init weights' between reference and product."""
    import hashlib
    h = hashlib.sha256()
    for k in sd:
        t = sd[k].detach().cpu().contiguous()
        h.update(k.encode())
        h.update(str(tuple(t.shape)).encode())
        if t.is_complex():
            t = torch.view_as_real(t)
        h.update(t.numpy().tobytes())
    return h.hexdigest()
