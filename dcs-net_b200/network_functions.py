"""Drop-in replacements for the hot-path part of the reference's network_functions.py (same names / signatures):
ComplexLReLU, ComplexSigmoid, ComplexAdaptiveAvgPool2d, ComplexAdaptiveMaxPool2d, bound_cRM, cRM, complex_mat_mult,
mag_phase_2_wave (/root/reference/network_functions.py:62-150) and the dcs / dc combine step
(network_functions.py:393-401, 431-436) as `enhance_batch`.  All arithmetic runs in sm_100a kernels (CUDA tensors only).

SiSNR / wSDR (network_functions.py:30-60) are loss/metric code outside the forward path; they are
kept as plain tensor expressions because `config.Config` instantiates them and SiSNR is the |dSI-SDR| parity metric.
"""
import torch

from . import _lib as L
from . import ops
from .complexFunctions import to_cl, from_cl, _eltwise_act


# ------------------------------------------------------------------ metrics (not on the hot path)
class SiSNR(object):
    def __call__(self, clean, estimate, eps=1e-8):
        dot = torch.sum(estimate * clean, -1, keepdim=True)
        energy = torch.sum(clean * clean, -1, keepdim=True)
        target = dot * clean / (energy + eps)
        resid = estimate - target
        ratio = torch.sum(target * target, -1, keepdim=True) / (torch.sum(resid * resid, -1, keepdim=True) + eps)
        return torch.mean(10 * torch.log10(ratio + eps))


class wSDR(object):
    def __call__(self, mixed, clean, clean_est, eps=2e-8):
        def neg_cos(a, b):
            return -(torch.sum(a * b, dim=1) / (torch.norm(a, p=2, dim=1) * torch.norm(b, p=2, dim=1) + eps))
        noise, noise_est = mixed - clean, mixed - clean_est
        e_c, e_n = torch.sum(clean ** 2, dim=1), torch.sum(noise ** 2, dim=1)
        alpha = e_c / (e_c + e_n + eps)
        return torch.mean(alpha * neg_cos(clean, clean_est) + (1 - alpha) * neg_cos(noise, noise_est))


# ------------------------------------------------------------------ element-wise layers
class ComplexLReLU(torch.nn.Module):
    def forward(self, input):
        return complex_lrelu(input)


def complex_lrelu(input):
    return _eltwise_act(input, L.ACT_LRELU)


class ComplexSigmoid(torch.nn.Module):
    def forward(self, input):
        return complex_sigmoid(input)


def complex_sigmoid(input):
    return _eltwise_act(input, L.ACT_SIGMOID)


def _global_avg(input, output_size):
    if output_size not in (1, (1, 1)):
        raise NotImplementedError("dcsnet_b200: adaptive pooling to 1x1 only (c_network.py:56-57)")
    L.require_cuda(input)
    x = to_cl(input)
    B, H, W, Cn, _ = x.shape
    sums = torch.zeros(B, Cn, 2, dtype=torch.float32, device=x.device)
    ops.chan_pool(x, sums)
    aff = torch.tensor([1.0 / (H * W), 0.0, 0.0, 1.0 / (H * W), 0.0, 0.0], device=x.device).repeat(Cn, 1).contiguous()
    y = ops.cbn_apply(sums.view(B, 1, 1, Cn, 2), aff)
    return from_cl(y)


class ComplexAdaptiveAvgPool2d(torch.nn.Module):
    def __init__(self, output_size):
        super(ComplexAdaptiveAvgPool2d, self).__init__()
        self.output_size = output_size

    def forward(self, input):
        return complex_adaptive_avg_pool2d(input, output_size=self.output_size)


def complex_adaptive_avg_pool2d(input, output_size=1):
    return _global_avg(input, output_size)


class ComplexAdaptiveMaxPool2d(torch.nn.Module):
    """The reference's "max" pool calls adaptive_avg_pool2d (network_functions.py:135-138): it IS an average pool."""

    def __init__(self, output_size):
        super(ComplexAdaptiveMaxPool2d, self).__init__()
        self.output_size = output_size

    def forward(self, input):
        return complex_adaptive_max_pool2d(input, output_size=self.output_size)


def complex_adaptive_max_pool2d(input, output_size=1):
    return _global_avg(input, output_size)


# ------------------------------------------------------------------ mask functions
def cRM(S, Y, eps=1e-8):
    return ops.crm(S, Y, eps)


def bound_cRM(cRM, hparams):
    return ops.bound_crm(cRM, hparams['atan2_eps'], exact_polar=True)


def complex_mat_mult(A, B):
    return ops.cmul(A, B)


def mag_phase_2_wave(mag, phase, config):
    if (config.fft_size, config.hop_length, config.window_length, config.normalise_stft) != (512, 32, 512, True):
        raise NotImplementedError("dcsnet_b200.mag_phase_2_wave: config.py:72-77 STFT parameters only (512/32/hann/normalized)")
    squeeze = mag.dim() == 2
    if squeeze:
        mag, phase = mag[None], phase[None]
    out = ops.istft_mag_phase(mag.contiguous(), phase.contiguous())
    return out[0] if squeeze else out


def enhance_batch(net, noisy_data, variant="dcs"):
    """The inference part of test_batch_2_metric_loss (network_functions.py:393-401 dcs, 431-436 dc):
    net(noisy) -> bound_cRM -> complex_mat_mult [-> subtraction] -> mag_phase_2_wave, run as the fused kernel plan.
    Returns dict(predict_noise_mask, predict_noise_data, predict_clean_data, predict_clean_audio[, predict_noise_audio])."""
    plan = net.plan_for(noisy_data, variant)
    out = plan.enhance_spec(noisy_data)
    res = dict(predict_noise_mask=out["mask"], predict_clean_data=out["clean_spec"], predict_noise_data=out["noise_spec"])
    eps = net.hparams['atan2_eps']
    res["predict_clean_audio"] = ops.istft(out["clean_spec"], atan2_eps=eps, exact_polar=plan.exact)
    if out["noise_spec"] is not None:
        res["predict_noise_audio"] = ops.istft(out["noise_spec"], atan2_eps=eps, exact_polar=plan.exact)
    return res
