"""Generate tests/golden/frontend.pt with the REFERENCE'S OWN resampler (build container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden_frontend

`Config().resample` (config.py:61, imported unmodified through oracle/reference_harness.py) resamples seeded synthetic
48 kHz "files"; the remaining lines of VoiceBankDataset.__getitem__ (data.py:90-134: pad / crop at start_point /
noise = noisy - clean / torch.stft x3 with the config's own window and flags) are executed literally here, because
data.py itself cannot be imported (torchaudio.backend.sox_io_backend is gone in torchaudio 2.x and the VoiceBank files
are absent).  Cases: a long utterance with a random crop, an utterance shorter than the window (zero padding, start 0),
and a length that is not a multiple of 3.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frontend.pt")


def main():
    cfg = rh.load()["config"].Config()
    window = cfg.integer_win_size - cfg.hop_length               # data.py:91
    g = torch.Generator().manual_seed(4321)
    cases = []
    for n48, start in ((48000 * 2 + 1, 5000), (20000, 0), (30011, 1233)):
        clean48 = 0.1 * torch.randn(n48, generator=g)
        noisy48 = clean48 + 0.05 * torch.randn(n48, generator=g)
        clean, noisy = cfg.resample(clean48), cfg.resample(noisy48)      # data.py:87-88
        data_len = clean.shape[0]
        if window > data_len:                                            # data.py:98-101
            clean = torch.nn.functional.pad(clean, (0, window - data_len))
            noisy = torch.nn.functional.pad(noisy, (0, window - data_len))
            start = 0
        clean = clean[start:start + window]                              # data.py:106-107
        noisy = noisy[start:start + window]
        noise = noisy - clean                                            # data.py:108
        st = lambda x: torch.stft(x, n_fft=cfg.fft_size, hop_length=cfg.hop_length, win_length=cfg.window_length,   # noqa: E731
                                  window=cfg.window, return_complex=True,
                                  normalized=cfg.normalise_stft)[1:int(cfg.fft_size / 2) + 1, :]
        # inputs are reproducible from the seed (the consumer re-draws them in the same order); spectrograms are kept
        # for every 16th frame only (the fixture stays ~0.6 MB)
        cases.append(dict(n48=n48, start=start, clean_audio=clean, noisy_audio=noisy, noise_audio=noise,
                          clean=st(clean)[:, ::16].clone(), noisy=st(noisy)[:, ::16].clone(), noise=st(noise)[:, ::16].clone()))
    torch.save(dict(window=window, seed=4321, frame_step=16, resample_kernel=cfg.resample.kernel.reshape(-1).clone(),
                    cases=cases), OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
