"""Public inference API: noisy audio in, enhanced audio out (the call a user of train.py/test.py's model makes when
they only want the enhanced waveform).  Host buffers are pinned; each call does H2D -> one CUDA-graph replay of the
fused kernel plan (STFT -> C_NETWORK -> bound_cRM x2 -> mask/subtract -> iSTFT) -> D2H on the current stream.

Long-form audio: the reference defines no overlap/stitching — it only ever enhances one crop of
`integer_win_size - hop` samples (config.py:110-111, data.py:91-104) — so long audio is cut into independent windows
of `window` samples (T = window/32 + 1 frames, T % 8 == 0), the last one zero-padded, enhanced as a batch and
concatenated in order.  Multi-GPU = contiguous shards of the window list, one process per GPU, no
data-path collective.
"""
import torch

from . import _lib as L
from .engine import ForwardPlan, PackedNet
from .rengine import RealForwardPlan, PackedRealNet

HOP = 32
WINDOW_4S = 63968   # 32 * (2000 - 1): "4 s" utterance with T = 2000 frames (T % 8 == 0)
WINDOW_REF = 8160   # 32 * (256 - 1): the reference's native 0.51 s crop (config.py:110-111)


def frames_for(n_samples):
    if n_samples % HOP:
        raise ValueError(f"window length must be a multiple of hop={HOP} (got {n_samples})")
    T = n_samples // HOP + 1
    if T % 8:
        raise ValueError(f"window of {n_samples} samples gives T={T} frames; C_NETWORK needs T % 8 == 0 "
                         f"(use n_samples = 32*(8k-1), e.g. {WINDOW_REF} or {WINDOW_4S})")
    return T


def shard_range(n_items, rank, world):
    """Contiguous [start, stop) of `n_items` owned by `rank` out of `world` (ceil split, trailing ranks may be empty)."""
    per = (n_items + world - 1) // world
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def split_windows(audio_1d, window):
    """1-D waveform -> (n_windows, window) with the tail zero-padded; returns (windows, original_length)."""
    n = audio_1d.numel()
    n_win = max(1, (n + window - 1) // window)
    out = audio_1d.new_zeros(n_win * window)
    out[:n] = audio_1d
    return out.view(n_win, window), n


class Enhancer:
    """Fixed-shape enhancer: `batch` windows of `n_samples` samples per call."""
    _packed_cls, _plan_cls, _variants = PackedNet, ForwardPlan, ("dcs", "dc")

    def __init__(self, model_or_sd, batch, n_samples=WINDOW_4S, mode="fp16", variant="dcs", device=None, graph=True,
                 atan2_eps=10e-7, exact_polar=False):
        if not torch.cuda.is_available():
            raise RuntimeError("dcsnet_b200.Enhancer needs a CUDA device (sm_100a); there is no CPU fallback")
        L.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.batch, self.n_samples, self.T = batch, n_samples, frames_for(n_samples)
        if variant not in self._variants:
            raise ValueError(f"{type(self).__name__} serves variants {self._variants}, got {variant!r}")
        with torch.cuda.device(self.device):
            self.packed = model_or_sd if isinstance(model_or_sd, self._packed_cls) else self._packed_cls(model_or_sd, self.device, mode)
            self.plan = self._plan_cls(self.packed, batch, self.T, variant=variant, atan2_eps=atan2_eps,
                                       exact_polar=exact_polar, want_aux=False)
            if graph:
                self.plan.capture()
        self.host_in = torch.empty(batch, n_samples, dtype=torch.float32, pin_memory=True)
        self.host_out = torch.empty(batch, n_samples, dtype=torch.float32, pin_memory=True)
        self.h2d_bytes = self.host_in.numel() * 4
        self.d2h_bytes = self.host_out.numel() * 4
        self._stream_state = None

    def enhance_device(self, audio_dev=None):
        """Device-resident path: audio (batch, n_samples) on the GPU (or already in plan.audio_in) -> device tensor."""
        return self.plan.enhance_audio(audio_dev)

    def enhance_pinned(self):
        """host_in (pinned) -> H2D -> graph -> D2H -> host_out (pinned).  Asynchronous on the current stream of the
        enhancer's device."""
        with torch.cuda.device(self.device):
            self.plan.audio_in.copy_(self.host_in, non_blocking=True)
            self.plan.enhance_audio()
            self.host_out.copy_(self.plan.audio_out, non_blocking=True)
        return self.host_out

    # ------------------------------------------------------------------ streaming (copy / compute overlap)
    def _streaming(self):
        if self._stream_state is None:
            with torch.cuda.device(self.device):
                st = dict(
                    h2d=torch.cuda.Stream(), d2h=torch.cuda.Stream(), i=0,
                    dev_in=[torch.empty_like(self.plan.audio_in) for _ in range(2)],
                    dev_out=[torch.empty_like(self.plan.audio_out) for _ in range(2)],
                    in_ready=[torch.cuda.Event() for _ in range(2)], in_free=[torch.cuda.Event() for _ in range(2)],
                    out_ready=[torch.cuda.Event() for _ in range(2)], out_free=[torch.cuda.Event() for _ in range(2)])
                cur = torch.cuda.current_stream()
                for k in range(2):
                    st["in_free"][k].record(cur)
                    st["out_free"][k].record(cur)
            self._stream_state = st
        return self._stream_state

    def enhance_pinned_stream(self, src=None, dst=None):
        """Streaming variant of enhance_pinned(): the H2D copy of this call and the D2H copy of the previous one run on
        their own streams through double-buffered device staging, so in steady state a step costs max(compute, copies).
        `src` / `dst`: pinned host tensors of n <= batch windows (default host_in / host_out).  The host buffers are read /
        written asynchronously: `dst` holds the result of this call only after drain()."""
        with torch.cuda.device(self.device):
            return self._enhance_pinned_stream(self.host_in if src is None else src, self.host_out if dst is None else dst)

    def _enhance_pinned_stream(self, src, dst):
        st = self._streaming()
        k = st["i"] & 1
        st["i"] += 1
        n = src.shape[0]
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(st["h2d"]):
            st["h2d"].wait_event(st["in_free"][k])              # the step that last read this staging buffer has consumed it
            st["dev_in"][k][:n].copy_(src, non_blocking=True)
            st["in_ready"][k].record(st["h2d"])
        cur.wait_event(st["in_ready"][k])
        self.plan.audio_in.copy_(st["dev_in"][k], non_blocking=True)
        st["in_free"][k].record(cur)
        self.plan.enhance_audio()
        cur.wait_event(st["out_free"][k])
        st["dev_out"][k].copy_(self.plan.audio_out, non_blocking=True)
        st["out_ready"][k].record(cur)
        with torch.cuda.stream(st["d2h"]):
            st["d2h"].wait_event(st["out_ready"][k])
            dst.copy_(st["dev_out"][k][:dst.shape[0]], non_blocking=True)
            st["out_free"][k].record(st["d2h"])
        return dst

    def drain(self):
        """Make the current stream wait for every copy issued by enhance_pinned_stream()."""
        if self._stream_state is not None:
            cur = torch.cuda.current_stream(self.device)
            for k in range(2):   # waiting on a never-recorded event is a no-op
                cur.wait_event(self._stream_state["out_free"][k])
                cur.wait_event(self._stream_state["in_ready"][k])

    def __call__(self, noisy_audio):
        """noisy_audio: (n, n_samples) CPU or CUDA float32, n <= batch.  Returns enhanced audio on the same device."""
        n = noisy_audio.shape[0]
        if noisy_audio.shape[1] != self.n_samples or n > self.batch:
            raise ValueError(f"expected (<= {self.batch}, {self.n_samples}) audio, got {tuple(noisy_audio.shape)}")
        with torch.cuda.device(self.device):
            if noisy_audio.is_cuda:
                if n < self.batch:
                    self.plan.audio_in[n:].zero_()
                self.plan.audio_in[:n].copy_(noisy_audio)
                return self.plan.enhance_audio()[:n].clone()
            self.host_in[:n].copy_(noisy_audio)
            if n < self.batch:
                self.host_in[n:].zero_()
            self.enhance_pinned()
            torch.cuda.current_stream().synchronize()
            return self.host_out[:n].clone()

    def enhance_windows_stream(self, windows_host, out_host=None):
        """Long-form path: (n_windows, n_samples) PINNED host windows -> enhanced windows in `out_host` (pinned; allocated if
        None), `batch` windows per step through enhance_pinned_stream(): every step's H2D and D2H copies overlap the
        neighbouring steps' compute, nothing synchronises between steps.  Asynchronous: call drain() + a stream sync before
        reading `out_host`.  A short last step reuses the stale tail rows of the staging buffer (windows are independent
        and those outputs are not copied back)."""
        n = windows_host.shape[0]
        assert windows_host.shape[1] == self.n_samples and windows_host.dtype == torch.float32
        if out_host is None:
            out_host = torch.empty(n, self.n_samples, dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(self.device):
            for i in range(0, n, self.batch):
                self._enhance_pinned_stream(windows_host[i:i + self.batch], out_host[i:i + self.batch])
        return out_host

    def enhance_long(self, audio_1d):
        """Arbitrary-length 1-D waveform -> enhanced waveform of the same length (independent windows, in order), streamed
        through enhance_windows_stream()."""
        wins, n = split_windows(audio_1d.detach().float().cpu(), self.n_samples)
        wins = wins.pin_memory()
        with torch.cuda.device(self.device):
            out = self.enhance_windows_stream(wins)
            self.drain()
            torch.cuda.current_stream().synchronize()
        return out.reshape(-1)[:n].clone()


class RealEnhancer(Enhancer):
    """Enhancer for the real-valued DR-Net / DRS-Net (R_NETWORK weights; rengine.RealForwardPlan): same host API."""
    _packed_cls, _plan_cls, _variants = PackedRealNet, RealForwardPlan, ("dr", "drs")

    def __init__(self, model_or_sd, batch, n_samples=WINDOW_4S, mode="fp16", variant="drs", **kw):
        super().__init__(model_or_sd, batch, n_samples=n_samples, mode=mode, variant=variant, **kw)
