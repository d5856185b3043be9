"""Data front-end (data.py:68-143): oracle vs the reference-generated golden vectors (CPU) and the CUDA
front-end vs the oracle / golden vectors through the C ABI (GPU)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend_oracle as FO  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "frontend.pt")


def rel_err(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.is_complex():
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def golden_cases():
    g = torch.load(GOLDEN)
    gen = torch.Generator().manual_seed(g["seed"])
    for c in g["cases"]:                       # the inputs are re-drawn exactly as oracle/make_golden_frontend.py drew them
        clean48 = 0.1 * torch.randn(c["n48"], generator=gen)
        noisy48 = clean48 + 0.05 * torch.randn(c["n48"], generator=gen)
        yield g, c, clean48, noisy48


def test_oracle_taps_equal_reference_resampler():
    g = torch.load(GOLDEN)
    k, width, orig, new = FO.sinc_resample_kernel()
    assert (width, orig, new) == (19, 3, 1)
    assert torch.equal(k.reshape(-1), g["resample_kernel"])      # bit-exact against config.resample.kernel


def test_oracle_getitem_matches_reference_golden():
    for g, c, clean48, noisy48 in golden_cases():
        out = FO.getitem(clean48, noisy48, c["start"], g["window"])
        for k in ("clean_audio", "noisy_audio", "noise_audio"):
            assert torch.equal(out[k], c[k]), k                  # same conv1d, same order: bit-exact
        for k in ("clean", "noisy", "noise"):
            assert rel_err(out[k][:, ::g["frame_step"]], c[k]) <= 1e-6, k


def test_oracle_resample_matches_torchaudio_when_available():
    ta = pytest.importorskip("torchaudio")
    x = torch.randn(2, 40003, generator=torch.Generator().manual_seed(1))
    assert torch.equal(FO.resample(x), ta.transforms.Resample(48000, 16000)(x))


def test_start_point_rules_host_logic():
    """data.py:96-104: short utterances start at 0, long ones draw from [0, data_len - window), equal lengths fail as in
    the reference (torch.randint(0, 0))."""
    import dcsnet_b200 as D
    sp = D.GpuFrontEnd.draw_start_points([20000, 48000 * 2 + 1], 8160, torch.Generator().manual_seed(0))
    assert sp[0] == 0 and 0 <= sp[1] < 32001 - 8160
    with pytest.raises(ValueError):
        D.GpuFrontEnd.draw_start_points([8160 * 3], 8160)
    k, width, orig = D.frontend.sinc_resample_kernel()
    assert torch.equal(k.reshape(-1), torch.load(GOLDEN)["resample_kernel"]) and (width, orig) == (19, 3)


@pytest.mark.gpu
def test_gpu_frontend_matches_reference_golden_and_oracle():
    import dcsnet_b200 as D
    cases = list(golden_cases())
    g = cases[0][0]
    L48 = max(c["n48"] for _, c, _, _ in cases)
    B = len(cases)
    clean48 = torch.zeros(B, L48)
    noisy48 = torch.zeros(B, L48)
    for i, (_, c, a, b) in enumerate(cases):
        clean48[i, :c["n48"]], noisy48[i, :c["n48"]] = a, b
        noisy48[i, c["n48"]:] = 7.0                              # garbage behind the valid length must not leak in
    fe = D.GpuFrontEnd(window=g["window"])
    n0 = D._lib.launch_count()
    out = fe.prepare(clean48.cuda(), noisy48.cuda(), lengths48=[c["n48"] for _, c, _, _ in cases],
                     start_points=[c["start"] for _, c, _, _ in cases])
    torch.cuda.synchronize()
    assert D._lib.launch_count() - n0 == 4                       # 1 front-end + 3 STFT kernels
    for i, (_, c, a, b) in enumerate(cases):
        for k in ("clean_audio", "noisy_audio", "noise_audio"):
            assert rel_err(out[k][i], c[k]) <= 2e-6, (i, k)      # fp32 FIR, different summation order than conv1d
        for k in ("clean", "noisy", "noise"):
            assert rel_err(out[k][i][:, ::g["frame_step"]], c[k]) <= 4e-6, (i, k)
    assert int(out["flags"].sum()) == 0


@pytest.mark.gpu
def test_gpu_frontend_flags_non_finite_audio_like_the_reference():
    import dcsnet_b200 as D
    x = 0.1 * torch.randn(2, 60000, generator=torch.Generator().manual_seed(3))
    y = x.clone()
    y[1, 30000] = float("inf")
    fe = D.GpuFrontEnd(window=8160)
    with pytest.raises(Exception, match="inf, neginf or nan in noisy audio"):
        fe.prepare(x.cuda(), y.cuda(), start_points=[0, 9000])
    out = fe.prepare(x.cuda(), y.cuda(), start_points=[0, 0], check=False)    # the window [0, 8160) misses sample 10000
    assert int(out["flags"].sum()) == 0


@pytest.mark.gpu
def test_gpu_frontend_full_size_properties():
    """64 utterances of 12 s at 48 kHz -> 4 s windows (the BASELINE batch): noise + clean == noisy exactly, linearity of the
    resampler, and agreement with the oracle on a subset."""
    import dcsnet_b200 as D
    B, L48 = 64, 48000 * 12
    g = torch.Generator().manual_seed(11)
    clean48 = 0.1 * torch.randn(B, L48, generator=g)
    noisy48 = clean48 + 0.05 * torch.randn(B, L48, generator=g)
    window = 63968
    fe = D.GpuFrontEnd(window=window)
    starts = fe.draw_start_points([L48] * B, window, torch.Generator().manual_seed(5))
    out = fe.prepare(clean48.cuda(), noisy48.cuda(), start_points=starts)
    assert out["noisy"].shape == (B, 256, 2000)
    assert torch.equal(out["noisy_audio"] - out["clean_audio"], out["noise_audio"])
    for i in (0, 17, 63):
        ref = FO.getitem(clean48[i], noisy48[i], starts[i], window)
        assert rel_err(out["clean_audio"][i], ref["clean_audio"]) <= 2e-6
        assert rel_err(out["noise"][i], ref["noise"]) <= 4e-6
    both = fe.prepare((clean48 + noisy48).cuda(), noisy48.cuda(), start_points=starts, check=False)
    assert rel_err(both["clean_audio"], out["clean_audio"] + out["noisy_audio"]) <= 2e-6
