"""Seeded synthetic utterances for the entry-point shims (train.py / test.py) and examples: there is no VoiceBank data in the image.
Noise-like "speech" with three sinusoids for spectral structure plus white noise — the same recipe the parity tests use, restated here
so that nothing on the product side imports the test oracle."""
import math

import torch

SAMPLE_RATE = 16000


def synthetic_audio(batch, length, seed=1234):
    """Returns (clean, noise, noisy) float32 tensors of shape (batch, length), noisy = clean + noise."""
    g = torch.Generator().manual_seed(seed)
    clean = 0.1 * torch.randn(batch, length, generator=g)
    noise = 0.05 * torch.randn(batch, length, generator=g)
    t = torch.arange(length, dtype=torch.float32) / SAMPLE_RATE
    for f0, a in ((220.0, 0.08), (1330.0, 0.05), (3100.0, 0.03)):
        clean = clean + a * torch.sin(2 * math.pi * f0 * t)[None, :]
    return clean, noise, clean + noise
