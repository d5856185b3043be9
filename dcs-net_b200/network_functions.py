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
    sums = ops.zero_(torch.empty(B, Cn, 2, dtype=torch.int64, device=x.device))
    ops.chan_pool(x, sums)
    return from_cl(ops.pool_mean(sums, H * W).view(B, 1, 1, Cn, 2))


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


# ------------------------------------------------------------------ evaluation step (test.py's per-batch function)
def _variant(self, variant):
    import sys
    v = variant or getattr(self, "variant", None) or (sys.argv[1] if len(sys.argv) > 1 else None)
    if v not in ("dcs", "drs", "dc", "dr"):
        raise ValueError("dcsnet_b200: variant must be one of dcs / drs / dc / dr (the reference reads sys.argv[1])")
    return v


def calc_loss(self, variant=None, **kwargs):
    """network_functions.py:168-208, argument names unchanged; `variant` replaces the reference's read of sys.argv[1]
    (used as the default when not given).  Note the reference's precedence at line 196: 1 - (speech_alpha * loss)."""
    v = _variant(self, variant)
    hp, cfg = self.hparams, self.config
    if v in ("dcs", "drs"):
        t = hp['noise_loss_type']
        wsdr = lambda: cfg.wSDR(kwargs['noisy_audio'], kwargs['noise_audio'], kwargs['predict_noise_audio'])
        l1_mask = lambda: cfg.L1(kwargs['target_noise_mask'], kwargs['predict_noise_mask'])
        l1_audio = lambda: cfg.L1(kwargs['noise_audio'], kwargs['predict_noise_audio'])
        if t == 0:
            orig = l1_mask()
        elif t == 1:
            orig = wsdr()
        elif t == 2:
            orig = l1_mask() + l1_audio()
        elif t == 3:
            orig = wsdr() + l1_audio()
        elif t == 4:
            orig = wsdr() + l1_mask()
        elif t == 5:
            tm, pm = kwargs['target_noise_mask'], kwargs['predict_noise_mask']
            if tm.is_complex():
                orig = wsdr() + cfg.mse(tm.real, pm.real) + cfg.mse(tm.imag, pm.imag)
            else:
                orig = wsdr() + cfg.mse(tm, pm)
        elif t == 6:
            orig = -cfg.SiSNR(kwargs['noise_audio'], kwargs['predict_noise_audio'])
        else:
            raise ValueError("noise_loss_type must be 0..6")
        noise_loss = 1 - hp['speech_alpha'] * orig
    if hp['speech_loss_type'] != 0:
        raise ValueError("speech_loss_type must be 0")
    speech_loss = hp['speech_alpha'] * (-cfg.SiSNR(kwargs['clean_audio'], kwargs['predict_clean_audio']))
    if v in ("dcs", "drs"):
        return noise_loss, speech_loss, noise_loss + speech_loss
    return speech_loss


def calc_metric(clean_audio, predict_audio, config, metric):
    """network_functions.py:152-166: mean of a per-utterance CPU metric (pesq / stoi are third-party CPU code)."""
    from math import isnan
    vals = []
    for i in range(predict_audio.shape[0]):
        m = metric(clean_audio[i, :].cpu().numpy(), predict_audio[i, :].cpu().numpy(), config.sr)
        if not isnan(m):
            vals.append(m)
    return float(sum(vals)) / max(len(vals), 1)


def _eval_batch(self, noise_data, noisy_data, clean_data, dtype, variant, metrics):
    """Shared body of val_ / test_batch_2_metric_loss (network_functions.py:282-361, 363-448): the tensor work on the GPU
    kernels — three iSTFTs (`dcs_istft_fwd`), target mask (`dcs_crm` + `dcs_bound_crm`, or sigmoid(|N| / |Y|) for the real
    path), network + mask combine + iSTFT (`enhance_batch` / `enhance_batch_real`) — then `calc_loss`."""
    v = _variant(self, variant)
    eps = self.hparams['atan2_eps']
    noise_audio = ops.istft(noise_data, atan2_eps=eps, exact_polar=True)
    noisy_audio = ops.istft(noisy_data, atan2_eps=eps, exact_polar=True)
    clean_audio = ops.istft(clean_data, atan2_eps=eps, exact_polar=True)

    two_mask = v in ("dcs", "drs")             # the reference branches on sys.argv[1] alone; dtype picks the arithmetic
    if dtype == "complex":
        out = enhance_batch(self, noisy_data, variant="dcs" if two_mask else "dc")
        if two_mask:
            target_noise_mask = bound_cRM(cRM(noise_data, noisy_data), self.hparams)
    elif dtype == "real":
        from .r_network import enhance_batch_real
        out = enhance_batch_real(self, noisy_data, variant="drs" if two_mask else "dr", atan2_eps=eps)
        if two_mask:
            noise_mag, _ = ops.mag_phase(noise_data, eps, want_phase=False)
            target_noise_mask = torch.sigmoid(noise_mag / out["noisy_mag"])
    else:
        raise ValueError("dtype must be 'real' or 'complex'")
    predict_clean_audio = out["predict_clean_audio"]

    def average(name):
        fn = (metrics or {}).get(name)
        if fn is None:
            try:
                fn = getattr(__import__("pypesq" if name == "pesq" else "pystoi"), name)
            except ImportError:
                return float("nan")
        return calc_metric(clean_audio, predict_clean_audio, self.config, fn)
    pesq_av, stoi_av = average("pesq"), average("stoi")

    if two_mask:
        predict_noise_audio = out["predict_noise_audio"]
        noise_loss, speech_loss, total = calc_loss(self, variant=v, target_noise_mask=target_noise_mask,
                                                   predict_noise_mask=out["predict_noise_mask"],
                                                   predict_noise_audio=predict_noise_audio,
                                                   predict_clean_audio=predict_clean_audio, noise_audio=noise_audio,
                                                   noisy_audio=noisy_audio, clean_audio=clean_audio)
        return (noise_loss, speech_loss, total, pesq_av, stoi_av, predict_noise_audio, predict_clean_audio,
                noise_audio, noisy_audio, clean_audio)
    speech_loss = calc_loss(self, variant=v, predict_clean_audio=predict_clean_audio, clean_audio=clean_audio)
    return speech_loss, pesq_av, stoi_av, predict_clean_audio, noise_audio, noisy_audio, clean_audio


def val_batch_2_metric_loss(self, val_batch, val_idx, dtype, variant=None, metrics=None):
    """network_functions.py:282-361, same return tuple.  `variant` replaces the reference's read of sys.argv[1];
    `metrics` = {"pesq": fn, "stoi": fn} (pypesq / pystoi signatures) — a metric that is neither given nor installed
    averages to NaN; nothing else on this path touches the CPU."""
    noise_data, noisy_data, clean_data, id = val_batch
    return _eval_batch(self, noise_data, noisy_data, clean_data, dtype, variant, metrics)


def test_batch_2_metric_loss(self, test_batch, test_idx, dtype, variant=None, metrics=None):
    """network_functions.py:363-448, same return tuple (dcs / drs append id and start_point; dc / dr do not, line 446)."""
    noise_data, noisy_data, clean_data, id, start_point = test_batch
    r = _eval_batch(self, noise_data, noisy_data, clean_data, dtype, variant, metrics)
    return r + (id, start_point) if len(r) == 10 else r


test_batch_2_metric_loss.__test__ = False   # not a pytest test


def epoch_end(self, outputs, type, variant=None):
    """network_functions.py:450-498: at the end of a validation / test epoch, log `config.val_log_sample_size` randomly chosen
    utterances (clean, predicted clean, noise, [predicted noise], noisy) to the experiment logger as audio.  Host-side glue of the
    Lightning layer: the draws go through numpy's global `random.choice` in the reference's order, so a seeded run
    logs the same samples; `variant` replaces sys.argv[1]."""
    from numpy import random
    two_mask = _variant(self, variant) in ("dcs", "drs")
    no_of_batches = len(outputs)
    random_batches = random.choice(no_of_batches, size=min(self.config.val_log_sample_size, no_of_batches), replace=False)
    keys = ['clean', 'predict_clean', 'noise'] + (['predict_noise'] if two_mask else []) + ['noisy']
    no_of_samples = min([self.config.data_params['batch_size']] + [outputs[-1][k].shape[0] for k in keys])
    random_samples = random.choice(no_of_samples, size=min(self.config.val_log_sample_size, no_of_samples), replace=False)
    for i, ridx in enumerate(range(min(self.config.val_log_sample_size, no_of_samples))):
        for k in keys:                                   # the reference's logging order: clean, predict_clean, noise, predict_noise, noisy
            self.logger.experiment.add_audio("{}({})/{}".format(k, type, i), outputs[random_batches[ridx]][k][random_samples[ridx], :],
                                             self.global_step, sample_rate=self.config.sr)
