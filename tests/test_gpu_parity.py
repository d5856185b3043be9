"""GPU (B200): parity of the CUDA path, called through the C ABI, against (i) the golden vectors produced by
executing the reference and (ii) the oracle restatement on the same seeded inputs.

Tolerances (north_star / SURVEY §8d), 'rel' = max|a-b| / max|b|:
  fp32 mode  : <= 1e-5 on the enhanced spectrogram S
  fp16 mode  : the tensor-core mode (tcgen05 kind::f16, fp16 storage, fp32 accumulation): <= 2e-3 on S and
               |dSI-SDR| <= 0.01 dB on EVERY weight state (default init and randomised BN), every golden (dcs, dc, B=1),
               per-layer taps <= 4e-3, and at the benchmark's own size (B=64 x T=2000)
  bf16 mode  : same kernels on bf16 storage (wide-range fallback): <= 2e-3 on default-init weights, <= 6e-3 on the
               randomised-BN state (8-bit significands; measured 3.6e-3) — documented, not the mode the bench is quoted on
  STFT/iSTFT : <= 2e-6 (fp32 FFT round-off)
Runs are bit-reproducible (fixed-point pooled sums): replays and batch permutations must be bit-identical.
"""
import pytest
import torch

import dcsnet_b200 as D
from dcsnet_b200 import ops, network_functions as NF, complexFunctions as CF, c_network as CN
from oracle import dcsnet_oracle as O, synthetic_weights as SW
from conftest import load_golden, rel_err, build_product_net

pytestmark = pytest.mark.gpu
EPS = O.HPARAMS["atan2_eps"]
TOL_FP32, TOL_TC, TOL_FFT = 1e-5, 2e-3, 2e-6
TOL_BF16_RANDBN = 6e-3
H16 = {"fp16": torch.float16, "bf16": torch.bfloat16}
ULP = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}     # one unit in the last place of the largest value


def si_sdr_delta(clean, a, b):
    return abs(float(O.si_snr(clean, a.cpu()) - O.si_snr(clean, b.cpu())))


def cl_to_nchw(t):
    return torch.view_as_complex(t.detach().float().cpu().contiguous()).permute(0, 3, 1, 2)


# ------------------------------------------------------------------ front / back end
def test_library_is_loaded_and_counts_launches():
    before = D._lib.launch_count()
    ops.stft(torch.zeros(1, 512, device="cuda"))
    torch.cuda.synchronize()
    assert D._lib.launch_count() == before + 1


@pytest.mark.parametrize("key", ["a", "b", "c"])
def test_stft_golden(key):
    c = load_golden("stft_istft.pt")[key]
    got = ops.stft(c["audio"].cuda())
    assert got.shape == c["spec"].shape
    assert rel_err(got, c["spec"]) <= TOL_FFT


@pytest.mark.parametrize("key", ["a", "b", "c"])
@pytest.mark.parametrize("exact", [False, True])
def test_istft_golden(key, exact):
    c = load_golden("stft_istft.pt")[key]
    got = ops.istft(c["ispec"].cuda(), atan2_eps=EPS, exact_polar=exact)
    assert got.shape == c["iwave"].shape
    assert rel_err(got, c["iwave"]) <= TOL_FFT


@pytest.mark.parametrize("B,T", [(1, 17), (3, 256), (2, 2000)])
def test_stft_istft_vs_oracle_ragged_and_full_length(B, T):
    _, _, noisy = O.synthetic_audio(B, 32 * (T - 1), seed=T)
    spec = O.stft(noisy).contiguous()
    assert rel_err(ops.stft(noisy.cuda()), spec) <= TOL_FFT
    assert rel_err(ops.istft(spec.cuda(), atan2_eps=EPS), O.spec_to_wave(spec)) <= TOL_FFT


def test_mag_phase_2_wave_dropin():
    from dcsnet_b200 import config as cfg
    g = torch.Generator().manual_seed(1)
    s = torch.complex(torch.randn(2, 256, 40, generator=g), torch.randn(2, 256, 40, generator=g))
    mag, ph = torch.abs(s), torch.atan2(s.imag, s.real + EPS)
    got = NF.mag_phase_2_wave(mag.cuda(), ph.cuda(), cfg.config)
    assert rel_err(got, O.spec_to_wave(s)) <= TOL_FFT


def test_stft_linearity_property_full_size():
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn(8, 63968, generator=g).cuda(), torch.randn(8, 63968, generator=g).cuda()
    lhs = ops.stft(a + b)
    rhs = ops.stft(a) + ops.stft(b)
    assert rel_err(lhs, rhs) <= 5e-6


# ------------------------------------------------------------------ the network, golden vectors from the reference
@pytest.mark.parametrize("name", ["cnet_dcs_default_B2_T64.pt", "cnet_dcs_randbn_B2_T64.pt",
                                  "cnet_dcs_randbn_B1_T32.pt", "cnet_dc_randbn_B2_T32.pt"])
def test_fp32_mode_matches_reference_golden(name):
    g = load_golden(name)
    net = build_product_net(g["state"])
    assert SW.state_dict_digest(net.state_dict()) == g["weights_sha256"]
    pk = D.PackedNet(net, "cuda", "fp32")
    plan = D.ForwardPlan(pk, g["B"], g["T"], variant=g["variant"])
    audio = plan.enhance_audio(g["noisy_audio"].cuda())
    torch.cuda.synchronize()
    assert rel_err(plan.clean_spec, g["clean_spec"]) <= TOL_FP32
    net_out = plan.net_out.squeeze(0) if g["B"] == 1 else plan.net_out
    assert rel_err(net_out, g["net_out"]) <= TOL_FP32
    assert rel_err(audio, g["clean_audio"]) <= TOL_FP32


@pytest.mark.parametrize("name", ["cnet_dcs_default_B2_T64.pt", "cnet_dcs_randbn_B2_T64.pt",
                                  "cnet_dcs_randbn_B1_T32.pt", "cnet_dc_randbn_B2_T32.pt"])
@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_tensor_core_modes_match_reference_golden(name, mode):
    """The tensor-core mode against vectors produced by EXECUTING the reference: default-init and randomised-BN weights,
    dcs and dc, B = 1 squeeze.  fp16 (the mode the bench is quoted on) meets the north-star 2e-3 / 0.01 dB everywhere."""
    g = load_golden(name)
    net = build_product_net(g["state"])
    plan = D.ForwardPlan(D.PackedNet(net, "cuda", mode), g["B"], g["T"], variant=g["variant"])
    audio = plan.enhance_audio(g["noisy_audio"].cuda())
    torch.cuda.synchronize()
    tol = TOL_TC if (mode == "fp16" or g["state"] == "default") else TOL_BF16_RANDBN
    assert rel_err(plan.clean_spec, g["clean_spec"]) <= tol
    assert rel_err(audio, g["clean_audio"]) <= tol
    clean = O.synthetic_audio(g["B"], 32 * (g["T"] - 1))[0]
    assert si_sdr_delta(clean, audio, g["clean_audio"]) <= 0.01


def test_module_forward_dropin_api():
    """C_NETWORK.forward(x) (the call c_network.py:187 serves): same shape conventions incl. the B=1 squeeze."""
    g = load_golden("cnet_dcs_randbn_B1_T32.pt")
    net = build_product_net("randbn").cuda()
    spec = O.stft(g["noisy_audio"])
    out = net(spec.cuda())
    assert out.shape == g["net_out"].shape == (256, 32)
    assert rel_err(out, g["net_out"]) <= TOL_FP32
    r = NF.enhance_batch(net, spec.cuda(), "dcs")
    assert rel_err(r["predict_clean_data"], g["clean_spec"]) <= TOL_FP32
    assert rel_err(r["predict_clean_audio"], g["clean_audio"]) <= TOL_FP32


# ------------------------------------------------------------------ per-layer taps vs the oracle (randomised BN)
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("fp16", 4e-3), ("bf16", 2.5e-2)])
def test_layer_taps_vs_oracle(mode, tol):
    sd = SW.make_state_dict(0)
    B, T = 2, 64
    _, _, noisy = O.synthetic_audio(B, 32 * (T - 1))
    taps = {}
    O.enhance_spec(sd, O.stft(noisy), taps=taps)
    plan = D.ForwardPlan(D.PackedNet(sd, "cuda", mode), B, T, keep_taps=True)
    plan.enhance_audio(noisy.cuda())
    torch.cuda.synchronize()
    assert rel_err(cl_to_nchw(plan.bn0), taps["bn0"]) <= (4e-3 if mode == "bf16" else (5e-4 if mode == "fp16" else tol))
    for k, v in plan.taps.items():
        if k in ("lstm", "fc"):
            got = torch.view_as_complex(v.detach().float().cpu().contiguous()).reshape(B, -1, 128)
        else:
            got = cl_to_nchw(v)
        assert rel_err(got, taps[k]) <= tol, k


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_tensor_core_conv_equals_cuda_core_conv_on_same_operands(mode):
    """tcgen05 kernel vs the FFMA kernel fed the SAME 16-bit-rounded activations and weights: only the fp32 summation
    order differs, so agreement must be ~1e-5 (isolates descriptor / swizzle / phase / two-source logic)."""
    sd = SW.make_state_dict(1)
    dt = H16[mode]
    pk = D.PackedNet(sd, "cuda", mode)
    g = torch.Generator().manual_seed(5)
    B, T = 2, 64
    cases = [(pk.enc[i], (B, 256 >> i, max(T >> i, T // 8), pk.enc[i].cin), False) for i in range(1, 7)]
    cases += [(pk.dec[i], (B, 2 << i, T // 8 if i < 4 else (T // 8) << (i - 4), pk.dec[i].cin // 2), True) for i in range(7)]
    for p, s0, two in cases:
        x0 = torch.randn(*s0, 2, generator=g).cuda().to(dt)
        x1 = torch.randn(*s0, 2, generator=g).cuda().to(dt) if two else None
        oh, ow = ops.conv_out_hw(p, s0[1], s0[2])
        ref = torch.empty(B, oh, ow, p.cout, 2, device="cuda")
        got = torch.empty(B, oh, ow, p.cout, 2, device="cuda")
        w_keep = p.w_ffma
        K = p.ntaps * 2 * p.cin
        p.w_ffma = p.w_tc.float()[:, :, :K].reshape(p.phases, p.n_pad, p.ntaps, 2 * p.cin).permute(0, 2, 3, 1).contiguous()
        ops.cconv(p, x0, x1, ref, use_tc=False)
        p.w_ffma = w_keep
        # fp32 output from the tensor-core path only exists for the last decoder layer; compare 16-bit-rounded otherwise
        if p is pk.dec[6]:
            ops.cconv(p, x0, x1, got, use_tc=True)
            tol = 2e-5
        else:
            gb = torch.empty(B, oh, ow, p.cout, 2, device="cuda", dtype=dt)
            ops.cconv(p, x0, x1, gb, use_tc=True)
            got, ref, tol = gb.float(), ref.to(dt).float(), ULP[dt]  # one ulp of the largest value
        torch.cuda.synchronize()
        assert not torch.isnan(got).any()
        assert rel_err(got, ref) <= tol, (p.cin, p.cout, p.up, p.stride)


@pytest.mark.parametrize("W", [250, 131])
@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_tensor_core_conv_wide_rows_pairs_and_dx_reuse(mode, W):
    """The general tcgen05 kernel at the geometry of the full-size step (image rows >= 128 pixels): tiles of 128 consecutive
    pixels of one row, CTA pairs (tcgen05.mma.cta_group::2, each CTA stages half of the weight rows; an odd tile count adds a
    masked tile) and dx-reuse staging (one A box per tap row, horizontal taps through row-shifted descriptors) vs the FFMA
    kernel on the SAME 16-bit activations and 16-bit-rounded weights; ragged width, odd batch, pooled sums."""
    sd = SW.make_state_dict(1)
    dt = H16[mode]
    pk = D.PackedNet(sd, "cuda", mode)
    g = torch.Generator().manual_seed(11)
    B = 3
    cases = [(pk.enc[i], (B, 8, W, pk.enc[i].cin), False) for i in range(3, 7)]
    cases += [(pk.dec[i], (B, 2 + i, W, pk.dec[i].cin // 2), True) for i in range(4)]
    for p, s0, two in cases:
        x0 = torch.randn(*s0, 2, generator=g).cuda().to(dt)
        x1 = torch.randn(*s0, 2, generator=g).cuda().to(dt) if two else None
        oh, ow = ops.conv_out_hw(p, s0[1], s0[2])
        ref = torch.empty(B, oh, ow, p.cout, 2, device="cuda")
        with rounded_ffma_weights(p):
            ops.cconv(p, x0, x1, ref, use_tc=False)
        gb = torch.full((B, oh, ow, p.cout, 2), float("nan"), device="cuda", dtype=dt)
        sums = torch.zeros(B, p.cout, 2, device="cuda", dtype=torch.int64)
        ops.cconv(p, x0, x1, gb, use_tc=True, pool_sums=sums)
        torch.cuda.synchronize()
        assert not torch.isnan(gb.float()).any(), (p.cin, p.cout)
        assert rel_err(gb.float(), ref.to(dt).float()) <= ULP[dt], (p.cin, p.cout, p.up, p.stride)
        want = ref.double().sum(dim=(1, 2))            # the epilogue sums its fp32 values (before the 16-bit rounding)
        assert rel_err(ops.pool_sums_to_float(sums), want) <= 1e-4, (p.cin, p.cout)


class rounded_ffma_weights:
    """Give the CUDA-core reference kernel the SAME 16-bit-rounded weights the tensor-core kernel multiplies (the packed
    tcgen05 operand, re-laid-out), so that a comparison isolates the data movement: a wrong small tap cannot hide inside a
    weight-rounding tolerance."""

    def __init__(self, p):
        self.p = p

    def __enter__(self):
        p = self.p
        K = p.ntaps * 2 * p.cin
        w = p.w_tc.float()[:, :, :K].reshape(p.phases, p.n_pad, p.ntaps, 2 * p.cin)          # [p][n][t][k]
        self.keep = (p.w_ffma, p.w_tail)
        p.w_ffma = w.permute(0, 2, 3, 1).contiguous()                                          # [p][t][k][n]
        if p.w_tail is not None:
            p.w_tail = w.permute(0, 2, 1, 3)[:, :, :2].reshape(p.phases, p.ntaps, 2, p.cin, 2).permute(0, 1, 3, 4, 2).contiguous()
        return p

    def __exit__(self, *a):
        self.p.w_ffma, self.p.w_tail = self.keep


@pytest.mark.parametrize("layer,merged,groups", [("enc1", True, 1), ("enc2", True, 1), ("dec5", True, 1), ("dec4", False, 2), ("dec4", True, 2)])
@pytest.mark.parametrize("B,H,W", [(2, 8, 66), (3, 20, 300)])
@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_strip_conv_equals_cuda_core_conv_on_same_operands(layer, merged, groups, B, H, W, mode):
    """Row-strip tcgen05 kernel (ring of source rows, row-shifted descriptors, resident weights) vs the FFMA kernel on
    the SAME 16-bit activations AND the same 16-bit-rounded weights: <= 1 ulp of the stored type; ragged widths (partial
    strips), several row chunks, fused (fixed-point) pooling sums."""
    from dcsnet_b200 import packing
    sd = SW.make_state_dict(1)
    dt = H16[mode]
    pk = D.PackedNet(sd, "cuda", mode)
    p = {"enc1": pk.enc[1], "enc2": pk.enc[2], "dec4": pk.dec[4], "dec5": pk.dec[5]}[layer]
    g = torch.Generator().manual_seed(7)
    if layer in ("enc1", "enc2"):
        c0, c1 = p.cin, 0
        x0, x1 = torch.randn(B, H, 2 * W, c0, 2, generator=g).cuda().to(dt), None
    else:
        c0 = c1 = p.cin // 2
        x0 = torch.randn(B, H, W, c0, 2, generator=g).cuda().to(dt)
        x1 = torch.randn(B, H, W, c1, 2, generator=g).cuda().to(dt)
    sp = packing.StripConv(p, c0, c1, merged=merged, groups=groups, device="cuda")
    oh, ow = ops.conv_out_hw(p, x0.shape[1], x0.shape[2])
    ref = torch.empty(B, oh, ow, p.cout, 2, device="cuda")
    with rounded_ffma_weights(p):
        ops.cconv(p, x0, x1, ref, use_tc=False)
    got = torch.full((B, oh, ow, p.cout, 2), float("nan"), device="cuda", dtype=dt)
    pool = torch.zeros(B, p.cout, 2, device="cuda", dtype=torch.int64)
    ops.cconv_strip(sp, x0, x1, got, pool_sums=pool)
    torch.cuda.synchronize()
    assert not torch.isnan(got.float()).any()
    assert rel_err(got.float(), ref.to(dt).float()) <= ULP[dt]
    assert rel_err(ops.pool_sums_to_float(pool), ref.double().sum(dim=(1, 2))) <= 1e-4   # the pool sums the UNROUNDED epilogue values
    pool2 = torch.zeros_like(pool)
    got2 = torch.empty_like(got)
    ops.cconv_strip(sp, x0, x1, got2, pool_sums=pool2)
    torch.cuda.synchronize()
    assert torch.equal(pool, pool2) and torch.equal(got.view(torch.int16), got2.view(torch.int16))   # bit-reproducible


@pytest.mark.parametrize("B,F,T", [(2, 16, 48), (3, 256, 272)])
@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_strip_enc0_equals_cuda_core_conv(B, F, T, mode):
    """encoder[0] on the tensor cores (Toeplitz blocks over 16-pixel strip rows) vs the FFMA conv on the same 16-bit
    initial_batchnorm output and the same rounded weights, with the fused pooling sums."""
    from dcsnet_b200 import packing
    sd = SW.make_state_dict(1)
    dt = H16[mode]
    pk = D.PackedNet(sd, "cuda", mode)
    p = pk.enc[0]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, F, T, 1, 2, generator=g).cuda().to(dt)
    sp = packing.StripEnc0(p, device="cuda")
    oh, ow = ops.conv_out_hw(p, F, T)
    ref = torch.empty(B, oh, ow, p.cout, 2, device="cuda")
    with rounded_ffma_weights(p):
        ops.cconv(p, x, None, ref, use_tc=False)
    got = torch.full((B, oh, ow, p.cout, 2), float("nan"), device="cuda", dtype=dt)
    pool = torch.zeros(B, p.cout, 2, device="cuda", dtype=torch.int64)
    ops.cconv_strip(sp, packing.StripEnc0.view_src(x), None, got, pool_sums=pool)
    torch.cuda.synchronize()
    assert not torch.isnan(got.float()).any()
    assert rel_err(got.float(), ref.to(dt).float()) <= ULP[dt]
    assert rel_err(ops.pool_sums_to_float(pool), ref.double().sum(dim=(1, 2))) <= 1e-4


@pytest.mark.parametrize("B,H,W,variant", [(2, 8, 72, "dcs"), (3, 128, 500, "dcs"), (2, 16, 40, "dc")])
@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_strip_dec6_tail_equals_cuda_core_tail(B, H, W, variant, mode):
    """decoder[6] + bound_cRM x2 + combine on the tensor cores (Toeplitz blocks over 4-pixel strip rows, tail epilogue)
    vs the FFMA dec6_tail kernel on the same 16-bit activations and the same rounded weights; every optional output."""
    from dcsnet_b200 import packing, _lib as L
    sd = SW.make_state_dict(1)
    dt = H16[mode]
    pk = D.PackedNet(sd, "cuda", mode)
    p = pk.dec[6]
    g = torch.Generator().manual_seed(13)
    d = torch.randn(B, H, W, 8, 2, generator=g).cuda().to(dt)
    k = torch.randn(B, H, W, 8, 2, generator=g).cuda().to(dt)
    Y = torch.complex(torch.randn(B, 2 * H, 2 * W, generator=g), torch.randn(B, 2 * H, 2 * W, generator=g)).cuda()
    combine = L.COMBINE_DCS if variant == "dcs" else L.COMBINE_DC
    new = lambda: torch.full((B, 2 * H, 2 * W), float("nan"), dtype=torch.complex64, device="cuda")
    ref = {n: new() for n in ("clean", "raw", "net", "mask", "noise")}
    got = {n: new() for n in ("clean", "raw", "net", "mask", "noise")}
    with rounded_ffma_weights(p):
        ops.dec6_tail(p, d, k, Y, ref["clean"], net_raw=ref["raw"], net_out=ref["net"], mask=ref["mask"], noise_spec=ref["noise"],
                      combine=combine)
    sp = packing.StripDec6(p, device="cuda")
    ops.dec6_tail_strip(sp, d, k, Y, got["clean"], net_raw=got["raw"], net_out=got["net"], mask=got["mask"],
                        noise_spec=got["noise"], combine=combine)
    torch.cuda.synchronize()
    for n in ref:
        if n == "noise" and variant == "dc":
            continue
        assert not torch.isnan(torch.view_as_real(got[n])).any(), n
        # identical operands: only the fp32 summation order (raw) and the MUFU-approximate bound_cRM (the rest) differ
        assert rel_err(got[n], ref[n]) <= (2e-5 if n == "raw" else 2e-4), n


@pytest.mark.parametrize("C,H,W", [(128, 2, 70), (128, 8, 37), (64, 16, 50), (32, 32, 33), (16, 64, 75), (8, 128, 130)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_fused_attention_equals_separate_kernels(C, H, W, dtype):
    """dcs_attention_fused (gate MLP + statistics + 7x7 gate conv + product in one pass, tile + halo in shared memory)
    vs the chan_pool -> chan_gate -> spat_stats -> spat_apply sequence; ragged widths, every layer geometry."""
    from dcsnet_b200 import packing
    sd = SW.make_state_dict(1)
    i = {128: 0, 64: 3, 32: 4, 16: 5, 8: 6}[C]
    ca = packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", "cuda")
    w7 = packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", "cuda")
    g = torch.Generator().manual_seed(C + H)
    B = 2
    x = torch.randn(B, H, W, C, 2, generator=g).cuda().to(dtype)
    sums = torch.zeros(B, C, 2, device="cuda", dtype=torch.int64)
    ops.chan_pool(x, sums)
    assert rel_err(ops.pool_sums_to_float(sums), x.double().sum(dim=(1, 2))) <= 1e-5
    gate = torch.empty(B, C, 2, device="cuda")
    stats = torch.empty(B, H * W, 4, device="cuda")
    ref = torch.empty_like(x)
    ops.chan_gate(sums, H * W, ca, gate)
    ops.spat_stats(x, gate, stats)
    ops.spat_apply(x, gate, stats, w7, ref)
    got = torch.full_like(x, float("nan"))
    ops.attention_fused(x, sums, ca, w7, got)
    torch.cuda.synchronize()
    assert not torch.isnan(got.float()).any()
    assert rel_err(got.float(), ref.float()) <= (2e-6 if dtype == torch.float32 else ULP[dtype])


@pytest.mark.parametrize("C,H,W", [(128, 2, 70), (128, 4, 33), (128, 8, 37), (64, 16, 50), (32, 32, 33), (16, 64, 75),
                                   (16, 64, 260), (8, 128, 130), (8, 128, 300), (8, 11, 16), (8, 30, 70), (16, 18, 50), (8, 9, 250)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_stream_attention_equals_separate_kernels(C, H, W, dtype):
    """dcs_attention_stream (16-bit storage: x rows through a bulk-copy ring, 7x7 gate conv as TF32 mma.sync row-partials with
    a register ring of pending rows) vs the chan_gate -> spat_stats -> spat_apply sequence with an fp32 output: every
    layer geometry, ragged widths, several column strips, heights that are not a multiple of the 4-row unroll."""
    from dcsnet_b200 import packing
    sd = SW.make_state_dict(1)
    i = {128: 0, 64: 3, 32: 4, 16: 5, 8: 6}[C]
    ca = packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", "cuda")
    w7 = packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", "cuda")
    g = torch.Generator().manual_seed(C + H + W)
    B = 3
    x = torch.randn(B, H, W, C, 2, generator=g).cuda().to(dtype)
    sums = torch.zeros(B, C, 2, device="cuda", dtype=torch.int64)
    ops.chan_pool(x, sums)
    gate = torch.empty(B, C, 2, device="cuda")
    stats = torch.empty(B, H * W, 4, device="cuda")
    ref = torch.empty(B, H, W, C, 2, device="cuda")
    ops.chan_gate(sums, H * W, ca, gate)
    ops.spat_stats(x, gate, stats)
    ops.spat_apply(x, gate, stats, w7, ref)
    got = torch.full_like(x, float("nan"))
    ops.attention_stream(x, sums, ca, w7, got)
    torch.cuda.synchronize()
    assert not torch.isnan(got.float()).any()
    # output rounding (half an ulp of the largest element: 2^-11 / 2^-8) + TF32 gate conv (~1e-4)
    assert rel_err(got.float(), ref) <= (8e-4 if dtype == torch.float16 else 5e-3)
    got2 = torch.empty_like(got)
    ops.attention_stream(x, sums, ca, w7, got2)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), got2.view(torch.int16))


# ------------------------------------------------------------------ properties at BASELINE size (B=64 x 4 s)
@pytest.mark.parametrize("mode,tol", [("fp32", TOL_FP32), ("fp16", TOL_TC)])
@pytest.mark.parametrize("state", ["default", "randbn"])
def test_full_size_batch_subset_vs_oracle_and_batch_independence(mode, tol, state):
    """The benchmark's own configuration (batch 64 x 3.998 s; default-init = the bench's weights, randbn = the survey's
    parity state): direct parity of utterances of the full-size batch, |dSI-SDR|, batch independence and reproducibility."""
    sd = {k: v.detach() for k, v in build_product_net(state).state_dict().items()}
    B, T = 64, 2000
    clean, _, noisy = O.synthetic_audio(B, 32 * (T - 1))
    pk = D.PackedNet(sd, "cuda", mode)
    plan = D.ForwardPlan(pk, B, T, want_aux=False).capture()
    full = plan.enhance_audio(noisy.cuda()).clone()
    spec_full = plan.clean_spec.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    # (i) direct parity of two utterances of the full-size batch against the oracle
    idx = [0, 63]
    ref = O.enhance_audio(sd, noisy[idx])
    assert rel_err(spec_full[idx], ref["clean_spec"]) <= tol
    assert rel_err(full[idx], ref["clean_audio"]) <= tol
    assert si_sdr_delta(clean[idx], full[idx], ref["clean_audio"]) <= 0.01
    # (ii) two replays of the same graph are bit-identical (integer pooled sums: no float atomics anywhere on the path)
    again = plan.enhance_audio().clone()
    torch.cuda.synchronize()
    assert torch.equal(again, full)
    # (iii) the reference's own CheckBatchGradient idea (network_functions.py:517-532): samples are independent —
    # permuting the batch permutes the output, bit for bit
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    out_p = plan.enhance_audio(noisy[perm].cuda())
    torch.cuda.synchronize()
    assert torch.equal(out_p, full[perm.cuda()])
    del plan


# ------------------------------------------------------------------ layer-wise drop-in classes
def test_layer_classes_match_oracle_functions():
    net = build_product_net("randbn").cuda()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    rc = lambda *s: torch.complex(torch.randn(*s, generator=g), torch.randn(*s, generator=g))
    x = rc(2, 16, 32, 24)
    # ComplexConv2d + ComplexBatchNorm2d + ComplexReLU (encoder block 2)
    ref = O.crelu(O.cbn_eval(O.cconv2d(x, sd, "encoder.2.0.", (2, 2), 2), sd, "encoder.2.1."))
    assert rel_err(net.encoder[2](x.cuda()), ref) <= 1e-5
    # ComplexConvTranspose2d (decoder 5 conv) on an explicit cat+upsample input
    u = O.cupsample_nearest(rc(2, 32, 8, 6), (2, 2))
    v = rc(1, 4, 3, 5)
    assert torch.equal(CF.complex_upsample(v.cuda(), scale_factor=(2, 1)).cpu(), O.cupsample_nearest(v, (2, 1)))
    assert rel_err(net.decoder[5][0](u.cuda()), O.cconvT2d(u, sd, "decoder.5.0.", 1, 1)) <= 1e-5
    # attention gates
    y = rc(2, 64, 16, 12)
    assert rel_err(net.skip_attention[6](y.cuda()), O.channel_attention(y, sd, "skip_attention.6.")) <= 1e-5
    assert rel_err(net.skip_attention[7](y.cuda()), O.spatial_attention(y, sd, "skip_attention.7.")) <= 1e-5
    # ComplexLSTM, ComplexLinear
    z = rc(2, 12, 128)
    assert rel_err(net.lstm(z.cuda()), O.complex_lstm(z, sd, "lstm.", explicit=True)) <= 1e-5
    assert rel_err(net.fc(z.cuda()), O.clinear(z, sd, "fc.")) <= 1e-5
    # element-wise functions of network_functions.py
    m = rc(2, 256, 16)
    assert rel_err(NF.bound_cRM(m.cuda(), net.hparams), O.bound_crm(m)) <= 2e-6
    assert rel_err(NF.complex_mat_mult(m.cuda(), m.flip(0).cuda()), m * m.flip(0)) <= 2e-6
    assert rel_err(NF.complex_sigmoid(m.cuda()), O.csigmoid(m)) <= 2e-6
    assert rel_err(NF.complex_lrelu(y.cuda()), O.clrelu(y)) <= 1e-7
    avg = torch.complex(y.real.mean((2, 3), keepdim=True), y.imag.mean((2, 3), keepdim=True))
    assert rel_err(NF.ComplexAdaptiveMaxPool2d(1)(y.cuda()), avg) <= 2e-6   # the "max" pool is an average pool


def test_shape_contract_errors():
    sd = SW.make_state_dict(0)
    pk = D.PackedNet(sd, "cuda", "fp32")
    with pytest.raises(ValueError):
        D.ForwardPlan(pk, 1, 260)       # T % 8 != 0
    with pytest.raises(ValueError):
        D.ForwardPlan(pk, 1, 64, n_bins=192)  # F % 128 != 0
    with pytest.raises(RuntimeError):
        ops.stft(torch.zeros(1, 100, device="cuda"))  # shorter than the reflect pad


def test_enhancer_public_api_and_graph():
    sd = SW.make_state_dict(0)
    _, _, noisy = O.synthetic_audio(3, 8160)
    ref = O.enhance_audio(sd, noisy)["clean_audio"]
    enh = D.Enhancer(sd, batch=4, n_samples=8160, mode="fp32")
    out = enh(noisy)                       # host in, host out, through the CUDA graph
    assert not out.is_cuda and rel_err(out, ref) <= 1e-5
    out2 = enh(noisy.cuda())
    assert out2.is_cuda and rel_err(out2, ref) <= 1e-5
    long = torch.cat([noisy[0], noisy[1], noisy[2][:1000]])
    got = enh.enhance_long(long)
    assert got.shape == long.shape
    assert rel_err(got[:16320], ref[:2].reshape(-1)) <= 1e-5


@pytest.mark.gpu
def test_enhancer_streaming_matches_serial():
    """enhance_pinned_stream (copies on their own streams, double-buffered staging) returns what enhance_pinned returns,
    step after step, with the input changing every step."""
    sd = SW.make_state_dict(0)
    enh = D.Enhancer(sd, batch=2, n_samples=8160, mode="fp16")
    outs_serial, outs_stream = [], []
    inputs = [O.synthetic_audio(2, 8160, seed=100 + i)[2] for i in range(4)]
    for x in inputs:
        enh.host_in.copy_(x)
        enh.enhance_pinned()
        torch.cuda.synchronize()
        outs_serial.append(enh.host_out.clone())
    for x in inputs:
        enh.host_in.copy_(x)
        enh.enhance_pinned_stream()
        enh.drain()
        torch.cuda.synchronize()       # host_in is rewritten next iteration: its H2D copy must have been consumed
        outs_stream.append(enh.host_out.clone())
    for a, b in zip(outs_serial, outs_stream):
        assert torch.equal(a, b)       # the path is bit-reproducible
    assert rel_err(outs_stream[0], outs_stream[1]) > 1e-2   # and the inputs did change from step to step
