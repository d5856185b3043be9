"""GPU data front-end — the per-item work of the reference's `VoiceBankDataset.__getitem__`
(/root/reference/data.py:68-143) for a whole batch on the device:

    config.resample (torchaudio Resample 48 kHz -> 16 kHz, config.py:61; data.py:87-88)
    length check / zero padding / random window crop (data.py:90-107)
    noise = noisy - clean (108), check_inf_neginf_nan x3 (110-112)
    torch.stft x3 -> (noise, noisy, clean) spectrograms, bins 1..256 (115-134)

One `dcs_frontend_fwd` launch (csrc/frontend.cu) + three `dcs_stft_fwd` launches; no CPU / ATen fallback.  The STFT
non-finite checks of data.py:136-138 are implied: the STFT of finite audio is finite.
"""
import math

import torch

from . import _lib as L
from . import ops
import ctypes as C

FILE_SR, SR = 48000, 16000                      # config.py:59-60
WINDOW_TRAIN = 8192 - 32                        # integer_win_size - hop_length (config.py:110-111, data.py:91)


def sinc_resample_kernel(orig_freq=FILE_SR, new_freq=SR, lowpass_filter_width=6, rolloff=0.99):
    """Taps of torchaudio.transforms.Resample(orig, new) with its defaults (sinc_interp_hann): float64 arithmetic, fp32
    result, as torchaudio 0.9 ... 2.x compute it.  Returns (kernel (new/gcd, taps) fp32, width, orig/gcd)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None, None] / new + idx
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * scale
    return kernels.to(torch.float32).reshape(new, -1), width, orig


class GpuFrontEnd:
    """prepare(clean48, noisy48, ...) -> dict(noise, noisy, clean: (B,256,T) complex64; *_audio: (B, window) fp32;
    start_points).  `window` = samples kept at 16 kHz (data.py:91: 8160 for the training crop; any multiple of 32
    that gives T % 8 == 0 works for the network)."""

    def __init__(self, window=WINDOW_TRAIN, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("dcsnet_b200.GpuFrontEnd needs a CUDA device (sm_100a); there is no CPU fallback")
        L.lib()
        k, self.width, self.orig = sinc_resample_kernel()
        assert k.shape[0] == 1, "48 kHz -> 16 kHz has a single output phase"
        self.kernel = k.reshape(-1).contiguous().to(device)
        self.window, self.device = int(window), device

    @staticmethod
    def draw_start_points(lengths48, window, generator=None):
        """data.py:96-104: start = 0 when the utterance is shorter than the window, else randint(0, data_len - window)."""
        out = []
        for n48 in lengths48:
            data_len = -(-int(n48) * SR // FILE_SR)                # ceil(new * length / orig)
            if window > data_len:
                out.append(0)
            else:
                if data_len == window:                             # the reference's torch.randint(0, 0) raises here
                    raise ValueError("utterance length equals the window: the reference's randint(0, 0) fails (data.py:103)")
                out.append(int(torch.randint(0, data_len - window, (1,), generator=generator)))
        return out

    def prepare(self, clean48, noisy48, lengths48=None, start_points=None, generator=None, check=True):
        """clean48 / noisy48: (B, L48) fp32 CUDA tensors (rows zero-padded to a common L48), lengths48: valid lengths."""
        L.require_cuda(clean48, noisy48)
        assert clean48.dtype == torch.float32 and clean48.shape == noisy48.shape and clean48.dim() == 2
        clean48, noisy48 = clean48.contiguous(), noisy48.contiguous()
        B, L48 = clean48.shape
        if lengths48 is None:
            lengths48 = [L48] * B
        if start_points is None:
            start_points = self.draw_start_points(lengths48, self.window, generator)
        len_t = torch.tensor([int(v) for v in lengths48], dtype=torch.int64, device=self.device)
        start_t = torch.tensor([int(v) for v in start_points], dtype=torch.int64, device=self.device)
        outs = [torch.empty(B, self.window, dtype=torch.float32, device=self.device) for _ in range(3)]
        flags = torch.zeros(B, dtype=torch.int32, device=self.device)
        p = L.FrontendParams(L.ptr(clean48), L.ptr(noisy48), L.ptr(len_t), L.ptr(start_t), B, L48, self.window,
                             L.ptr(self.kernel), self.kernel.numel(), self.orig, self.width,
                             L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(outs[2]), L.ptr(flags))
        L.check(L.lib().dcs_frontend_fwd(C.byref(p), L.stream_ptr()), "dcs_frontend_fwd")
        clean16, noisy16, noise16 = outs
        res = dict(clean_audio=clean16, noisy_audio=noisy16, noise_audio=noise16, start_points=list(start_points),
                   clean=ops.stft(clean16), noisy=ops.stft(noisy16), noise=ops.stft(noise16), flags=flags)
        if check:
            f = flags.cpu()
            for bit, what in ((1, "clean"), (2, "noisy"), (4, "noise")):
                if bool((f & bit).any()):
                    raise Exception(f"Found inf, neginf or nan in {what} audio!")   # check_inf_neginf_nan's message (data.py:110-112)
        return res
