"""Real network path (r_network.py): the CPU oracle against vectors produced by the reference itself,
and the drop-in R_NETWORK parameter container (state_dict keys / shapes / seed-0 weights identical to the reference)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rnet_oracle as RO, dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402
from oracle.make_golden_rnet import randomise_bn  # noqa: E402
import dcsnet_b200 as D  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "rnet_drs_randbn_B2_T64.pt")


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def product_net():
    from dcsnet_b200 import r_network, config as C
    return r_network.R_NETWORK(C.Config(), dict(C.hparams), 0).eval()


def test_rnetwork_state_dict_and_seed0_weights_identical_to_reference():
    g = torch.load(GOLDEN)
    net = product_net()
    sd = net.state_dict()
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()] == g["keys"]
    assert SW.state_dict_digest(sd) == g["digest_seed0"]            # bit-identical random-init weights


def test_rnet_oracle_matches_reference_golden():
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    sd = net.state_dict()
    spec = O.stft(g["noisy_audio"])
    taps = {}
    mask = RO.r_network_forward(sd, torch.abs(spec), taps=taps)
    assert rel_err(mask, g["mask"]) <= 2e-6
    for k, fp in g["taps"].items():
        assert tuple(taps[k].shape) == fp["shape"]
        assert abs(float(taps[k].abs().max()) - fp["max_abs"]) <= 1e-5 * fp["max_abs"]
        assert rel_err(taps[k].reshape(-1)[:32], fp["head"]) <= 1e-5
    for variant in ("drs", "dr"):
        out = RO.enhance_spec(sd, spec, variant)
        assert rel_err(out["clean_mag"], g[f"{variant}_clean_mag"]) <= 2e-6
        assert rel_err(out["clean_audio"], g[f"{variant}_clean_audio"]) <= 2e-6


def test_channel_attention_uses_only_the_max_pool_branch():
    """r_network.py:23-24: `out = avg_out_fc + max_out_fc` is overwritten by `out = max_out_fc`."""
    g = torch.Generator().manual_seed(0)
    sd = {"a.fc.0.weight": torch.randn(2, 16, 1, 1, generator=g), "a.fc.2.weight": torch.randn(16, 2, 1, 1, generator=g)}
    x = torch.randn(3, 16, 5, 7, generator=g)
    got = RO.channel_attention(x, sd, "a.")
    y = x.clone()
    flat = y.view(3, 16, -1)
    flat.scatter_(-1, flat.argmin(-1, keepdim=True), -100.0)        # changes every average, no maximum
    assert torch.equal(got, RO.channel_attention(y, sd, "a."))


def test_product_forward_fails_loudly_without_fallback():
    net = product_net()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(2, 256, 64))
    with pytest.raises(NotImplementedError):
        net.train()(torch.rand(2, 256, 64))


@pytest.mark.gpu
def test_gpu_rnetwork_forward_matches_reference_golden():
    """R_NETWORK.forward on the B200 (fp32 kernel sequence) vs the mask the reference's own R_NETWORK produced."""
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    net = net.cuda().eval()
    spec = O.stft(g["noisy_audio"])
    n0 = D._lib.launch_count()
    mask = net(torch.abs(spec).cuda())
    torch.cuda.synchronize()
    assert D._lib.launch_count() - n0 >= 7 + 7 + 13 * 4 + 6          # convs, attentions, LSTM: all from libdcsnet_sm100a.so
    assert tuple(mask.shape) == tuple(g["mask"].shape)
    assert rel_err(mask.cpu(), g["mask"]) <= 1e-5


# ------------------------------------------------------------------ real convs on the complex-conv operand layout (CPU emulation)
def _pairs(t):
    """(B, C, H, W) real NCHW -> channels-last pseudo-complex (B, H, W, ceil(C/2), 2): channel pairs (2c, 2c+1) = (re, im)."""
    B, C, H, W = t.shape
    if C % 2:
        t = torch.cat([t, torch.zeros(B, 1, H, W)], dim=1)
    return t.permute(0, 2, 3, 1).reshape(B, H, W, -1, 2).contiguous()


@pytest.mark.parametrize("layer", ["enc0", "enc3", "dec0", "dec4", "dec6"])
def test_real_conv_packing_reproduces_the_reference_layers(layer):
    """packing.packed_conv_from_real + real_bn_affine: a real Conv2d / (cat + nearest Upsample + ConvTranspose2d) with the
    BatchNorm2d and activation that follow, evaluated through the SAME operand definitions the complex-conv kernels use
    (tests/emulate.conv_geometry), equals the torch layer stack of r_network.py."""
    from dcsnet_b200 import packing, _lib as L
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emulate import conv_geometry
    import torch.nn.functional as F
    net = product_net()
    randomise_bn(net.state_dict(), 7)
    sd = net.state_dict()
    g = torch.Generator().manual_seed(5)
    if layer.startswith("enc"):
        i = int(layer[3:])
        conv, bn = net.encoder[i][0], net.encoder[i][1]
        x = torch.randn(2, conv.in_channels, 12, 10, generator=g)
        pk = packing.packed_conv_from_real(conv.weight, conv.bias, bn=packing.real_bn_affine(bn.weight, bn.bias, bn.running_mean, bn.running_var),
                                           stride=conv.stride, act=L.ACT_RELU)
        ref = F.relu(F.batch_norm(F.conv2d(x, conv.weight, conv.bias, stride=conv.stride, padding=conv.padding), bn.running_mean,
                                  bn.running_var, bn.weight, bn.bias, False, 0.0, 1e-5))
        oh, ow = ref.shape[2:]
        got = conv_geometry(pk, _pairs(x), None, (oh, ow))
    else:
        i = int(layer[3:])
        last = i == 6
        convt = net.decoder[i] if last else net.decoder[i][0]
        up = RO.UPSAMPLE[i]
        c = convt.in_channels // 2
        d, skip = torch.randn(2, c, 4, 5, generator=g), torch.randn(2, c, 4, 5, generator=g)
        bn = None if last else net.decoder[i][1]
        pk = packing.packed_conv_from_real(convt.weight, convt.bias, transposed=True, up=up,
                                           bn=None if last else packing.real_bn_affine(bn.weight, bn.bias, bn.running_mean, bn.running_var),
                                           act=L.ACT_NONE if last else L.ACT_LRELU)
        ref = F.conv_transpose2d(F.interpolate(torch.cat((d, skip), 1), scale_factor=up, mode="nearest"), convt.weight, convt.bias,
                                 stride=1, padding=1)
        if not last:
            ref = F.leaky_relu(F.batch_norm(ref, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, 1e-5))
        oh, ow = ref.shape[2:]
        got = conv_geometry(pk, _pairs(d), _pairs(skip), (oh, ow))
    got = got.reshape(2, oh, ow, -1)[..., :ref.shape[1]].permute(0, 3, 1, 2)
    assert rel_err(got, ref) <= 2e-6


def test_whole_rnetwork_through_the_packed_operands_matches_reference_golden():
    """packing.PackedRNet (BN folds, channel pairing, concat order, sub-pixel phases of the up-sampling, LSTM / CBAM weight
    layouts) evaluated by tests/emulate.rnet_dataflow reproduces the mask the reference's R_NETWORK produced."""
    from dcsnet_b200 import packing
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emulate import rnet_dataflow
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    pk = packing.PackedRNet(net.state_dict())
    spec = O.stft(g["noisy_audio"])
    mask = rnet_dataflow(pk, torch.abs(spec))
    assert rel_err(mask, g["mask"]) <= 1e-5


@pytest.mark.gpu
def test_gpu_real_encoder_and_decoder_convs_on_the_existing_kernels():
    """First GPU piece of the real path: R_NETWORK's BatchNorm'd magnitude, its 7 encoder convs (fp32 CUDA-core kernels) and
    a decoder stage (cat + up-sampling + ConvTranspose2d + BN + LeakyReLU; fp32 and the tcgen05 bf16 kernel) run on the
    EXISTING complex-conv kernels from packing.PackedRNet operands and match the oracle's layer taps."""
    from dcsnet_b200 import packing, ops
    import torch.nn.functional as F
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    sd = net.state_dict()
    pk = packing.PackedRNet(sd, device="cuda", want_bf16=True)
    spec = O.stft(g["noisy_audio"])
    mag = torch.abs(spec)
    taps = {}
    RO.r_network_forward(sd, mag, taps=taps)
    B, Fb, T = mag.shape
    x = torch.zeros(B, Fb, T, 1, 2, device="cuda")
    x[..., 0, 0] = mag.cuda()
    x = ops.cbn_apply(x, pk.bn0)
    H, W = Fb, T
    for i in range(7):
        H, W = ops.conv_out_hw(pk.enc[i], H, W)
        y = torch.empty(B, H, W, pk.enc[i].cout, 2, device="cuda")
        ops.cconv(pk.enc[i], x, None, y, use_tc=False)
        x = y
        got = y.reshape(B, H, W, -1).permute(0, 3, 1, 2).cpu()
        assert rel_err(got, taps[f"enc{i}"]) <= 1e-5, i
    # decoder stage 1 on random inputs: (cat + Upsample(2,1) + ConvTranspose2d + BN + LeakyReLU)
    gen = torch.Generator().manual_seed(9)
    convt, bn = net.decoder[1][0], net.decoder[1][1]
    c = convt.in_channels // 2
    d, skip = torch.randn(2, c, 4, 8, generator=gen), torch.randn(2, c, 4, 8, generator=gen)
    ref = F.leaky_relu(F.batch_norm(F.conv_transpose2d(F.interpolate(torch.cat((d, skip), 1), scale_factor=(2, 1), mode="nearest"),
                                                       convt.weight, convt.bias, stride=1, padding=1),
                                    bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, 1e-5))
    cl = lambda t: t.permute(0, 2, 3, 1).reshape(t.shape[0], t.shape[2], t.shape[3], -1, 2).contiguous().cuda()   # noqa: E731
    for dtype, use_tc, tol in ((torch.float32, False, 1e-5), (torch.bfloat16, True, 1.5e-2)):
        out = torch.empty(2, 8, 8, pk.dec[1].cout, 2, device="cuda", dtype=dtype)
        ops.cconv(pk.dec[1], cl(d).to(dtype), cl(skip).to(dtype), out, use_tc=use_tc)
        got = out.float().reshape(2, 8, 8, -1).permute(0, 3, 1, 2).cpu()
        assert rel_err(got, ref) <= tol, dtype


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,W", [(16, 12, 9), (64, 7, 20), (256, 2, 5), (32, 33, 70)])
def test_gpu_real_attention_matches_oracle(C, H, W):
    """dcs_real_attention_fwd (RealChannelAttention + RealSpatialAttention, r_network.py:8-40) vs the oracle, fp32 and bf16."""
    from dcsnet_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    R = max(C // 16, 1)
    sd = {"a.fc.0.weight": 0.3 * torch.randn(R, C, 1, 1, generator=g), "a.fc.2.weight": 0.5 * torch.randn(C, R, 1, 1, generator=g),
          "s.conv1.weight": 0.2 * torch.randn(1, 2, 7, 7, generator=g)}
    x = torch.randn(3, C, H, W, generator=g)
    u = RO.channel_attention(x, sd, "a.") * x
    ref = RO.spatial_attention(u, sd, "s.") * u
    att = dict(w1=sd["a.fc.0.weight"].flatten(1).contiguous().cuda(), w2=sd["a.fc.2.weight"].flatten(1).contiguous().cuda())
    w7 = sd["s.conv1.weight"].reshape(2, 49).contiguous().cuda()
    xcl = x.permute(0, 2, 3, 1).contiguous().cuda()
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1.5e-2)):
        got = ops.real_attention(xcl.to(dtype), att, w7)
        torch.cuda.synchronize()
        assert rel_err(got.float().permute(0, 3, 1, 2).cpu(), ref) <= tol, dtype


@pytest.mark.gpu
def test_gpu_real_lstm_matches_torch_lstm():
    """dcs_rlstm_fwd (nn.LSTM(256 -> 128, 2 layers, bidirectional) of r_network.py:70-74) vs torch.nn.LSTM on the CPU."""
    from dcsnet_b200 import packing, ops
    net = product_net()
    pk = packing.PackedRNet(net.state_dict(), device="cuda")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 21, 256, generator=g)
    with torch.no_grad():
        ref, _ = net.lstm(x)
    got = ops.rlstm(x.cuda(), pk.lstm_t)
    torch.cuda.synchronize()
    assert rel_err(got.cpu(), ref) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["drs", "dr"])
def test_gpu_real_step_matches_reference_golden(variant):
    """|Y| / noisy phase / mask / magnitude combine / mag_phase_2_wave of the dr and drs step functions, all on the GPU."""
    from dcsnet_b200 import r_network
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    net = net.cuda().eval()
    out = r_network.enhance_batch_real(net, O.stft(g["noisy_audio"]).cuda(), variant)
    torch.cuda.synchronize()
    assert rel_err(out["predict_clean_mag"].cpu(), g[f"{variant}_clean_mag"]) <= 1e-5
    assert rel_err(out["predict_clean_audio"].cpu(), g[f"{variant}_clean_audio"]) <= 2e-5


@pytest.mark.parametrize("layer,merged,groups", [("enc1", True, 1), ("enc2", True, 1), ("dec4", False, 2), ("dec5", True, 1)])
def test_real_layers_pack_for_the_row_strip_kernel(layer, merged, groups):
    """The bf16 / tensor-core real path will reuse the row-strip kernel: packing.StripConv built from a REAL layer's
    PackedConv describes the same convolution as its per-tap operands (CPU emulation of the kernel's item table)."""
    from dcsnet_b200 import packing, ops
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emulate import conv_geometry, strip_geometry
    net = product_net()
    randomise_bn(net.state_dict(), 7)
    pk = packing.PackedRNet(net.state_dict(), want_bf16=True)
    p = {"enc1": pk.enc[1], "enc2": pk.enc[2], "dec4": pk.dec[4], "dec5": pk.dec[5]}[layer]
    g = torch.Generator().manual_seed(13)
    if layer.startswith("enc"):
        H, W = 6, 260
        srcs, c0, c1 = [torch.randn(2, H, W, p.cin, 2, generator=g), None], p.cin, 0
    else:
        H, W = 3, 131
        c = p.cin // 2
        srcs, c0, c1 = [torch.randn(2, H, W, c, 2, generator=g), torch.randn(2, H, W, c, 2, generator=g)], c, c
    srcs = [None if t is None else t.to(torch.bfloat16).float() for t in srcs]
    sp = packing.StripConv(p, c0, c1, merged=merged, groups=groups)
    out_hw = ops.conv_out_hw(p, H, W)
    assert rel_err(strip_geometry(sp, srcs[0], srcs[1], out_hw), conv_geometry(p, srcs[0], srcs[1], out_hw, weights="tc")) <= 1e-2


# ------------------------------------------------------------------ the real path on the tensor cores (rengine.RealForwardPlan)
@pytest.mark.gpu
@pytest.mark.parametrize("mode,B,S", [("fp16", 3, 37), ("fp16", 8, 500), ("bf16", 5, 21)])
def test_gpu_real_lstm_tensor_core_matches_torch_lstm(mode, B, S):
    """dcs_rlstm_tc_fwd (kind::f16 projections + fp16 mma.sync recurrence, W_hh in registers) vs torch.nn.LSTM on the same
    16-bit-rounded input; batch sizes that are not a multiple of the 4 sequences a CTA owns; the full sequence length."""
    from dcsnet_b200 import ops
    net = product_net()
    dt = torch.float16 if mode == "fp16" else torch.bfloat16
    pk = D.PackedRealNet(net.state_dict(), "cuda", mode)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, S, 256, generator=g).to(dt)
    with torch.no_grad():
        ref, _ = net.lstm(x.float())
    y = torch.full((B, S, 256), float("nan"), dtype=dt, device="cuda")
    ws = torch.empty(ops.rlstm_tc_workspace_bytes(B, S), dtype=torch.uint8, device="cuda")
    ops.rlstm_tc(x.cuda(), pk.lstm_tc, y, ws)
    torch.cuda.synchronize()
    assert not torch.isnan(y.float()).any()
    # 16-bit weights / hidden state / output (|h| < 1): a few ulps of the storage type
    assert rel_err(y.float().cpu(), ref) <= (3e-3 if mode == "fp16" else 2.5e-2)
    y2 = torch.empty_like(y)
    ops.rlstm_tc(x.cuda(), pk.lstm_tc, y2, ws)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int16), y2.view(torch.int16))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["drs", "dr"])
def test_gpu_real_tensor_core_plan_matches_reference_golden(variant):
    """RealForwardPlan (fp16 storage, tcgen05 convs, fused sigmoid / magnitude / noisy-phase tail) from the AUDIO against
    what the reference's own r_network.py + step functions produced: mask, clean magnitude, waveform <= 2e-3, |dSI-SDR|."""
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    B, n = g["noisy_audio"].shape
    T = n // 32 + 1
    plan = D.RealForwardPlan(D.PackedRealNet(net.state_dict(), "cuda", "fp16"), B, T, variant=variant)
    n0 = D._lib.launch_count()
    audio = plan.enhance_audio(g["noisy_audio"].cuda()).clone()
    torch.cuda.synchronize()
    assert D._lib.launch_count() - n0 >= 30
    assert rel_err(plan.mask.cpu(), g["mask"]) <= 2e-3
    assert rel_err(plan.clean_spec.abs().cpu(), g[f"{variant}_clean_mag"]) <= 2e-3
    assert rel_err(audio.cpu(), g[f"{variant}_clean_audio"]) <= 2e-3
    clean = O.synthetic_audio(B, n)[0]
    assert abs(float(O.si_snr(clean, audio.cpu()) - O.si_snr(clean, g[f"{variant}_clean_audio"]))) <= 0.01
    again = plan.enhance_audio(g["noisy_audio"].cuda())
    torch.cuda.synchronize()
    assert torch.equal(again, audio)           # bit-reproducible


@pytest.mark.gpu
def test_gpu_real_tensor_core_plan_full_size_vs_oracle():
    """BASELINE configs[2] size (batch 64 x 3.998 s) through RealEnhancer (CUDA graph): utterances of the full batch against
    the oracle, replays bit-identical, samples independent."""
    net = product_net()
    randomise_bn(net.state_dict(), 7)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    B, T = 64, 2000
    clean, _, noisy = O.synthetic_audio(B, 32 * (T - 1))
    enh = D.RealEnhancer(sd, batch=B, n_samples=32 * (T - 1), mode="fp16", variant="drs")
    full = enh.enhance_device(noisy.cuda()).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    idx = [0, 63]
    ref = RO.enhance_spec(sd, O.stft(noisy[idx]), "drs")["clean_audio"]
    assert rel_err(full[idx].cpu(), ref) <= 2e-3
    assert abs(float(O.si_snr(clean[idx], full[idx].cpu()) - O.si_snr(clean[idx], ref))) <= 0.01
    again = enh.enhance_device().clone()
    torch.cuda.synchronize()
    assert torch.equal(again, full)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    out_p = enh.enhance_device(noisy[perm].cuda())
    torch.cuda.synchronize()
    assert torch.equal(out_p, full[perm.cuda()])


@pytest.mark.gpu
@pytest.mark.parametrize("Cp,H,W", [(128, 2, 70), (128, 8, 37), (64, 16, 50), (32, 32, 33), (16, 64, 260), (8, 128, 300), (8, 11, 16), (8, 30, 70), (16, 18, 50)])
def test_gpu_real_stream_attention_equals_three_pass_kernels(Cp, H, W):
    """The REAL variant of dcs_attention_stream (max-pool channel gate from DCS_POOL_MAX-encoded maxima, (mean, max)
    statistics, TF32 mma.sync gate conv, element-wise products) vs dcs_real_attention_fwd (fp32 output) and the oracle's
    torch functions; every layer geometry, ragged widths; stand-alone dcs_chan_max."""
    from dcsnet_b200 import ops, packing
    net = product_net()
    pk = packing.PackedRNet(net.state_dict(), device="cuda")
    i = {128: 0, 64: 3, 32: 4, 16: 5, 8: 6}[Cp]
    att, w7 = pk.skip_att[i]
    g = torch.Generator().manual_seed(Cp + H + W)
    B = 3
    x = torch.randn(B, H, W, Cp, 2, generator=g).to(torch.float16)
    maxima = torch.zeros(B, Cp, 2, dtype=torch.int64, device="cuda")
    ops.chan_max(x.cuda(), maxima)
    ref = ops.real_attention(x.float().cuda(), att, w7)
    got = torch.full((B, H, W, Cp, 2), float("nan"), dtype=torch.float16, device="cuda")
    ops.real_attention_stream(x.cuda(), maxima, att, w7, got)
    torch.cuda.synchronize()
    assert not torch.isnan(got.float()).any()
    assert rel_err(got.float().cpu(), ref.cpu()) <= 8e-4          # fp16 output rounding (2^-11) + TF32 gate conv (~1e-4)
    # and against the oracle's own functions on the same input (NCHW real view of the pair tensor)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    xr = x.float().reshape(B, H, W, 2 * Cp).permute(0, 3, 1, 2)
    u = xr * RO.channel_attention(xr, sd, f"skip_attention.{2 * i}.")
    want = (u * RO.spatial_attention(u, sd, f"skip_attention.{2 * i + 1}.")).permute(0, 2, 3, 1).reshape(B, H, W, Cp, 2)
    assert rel_err(got.float().cpu(), want) <= 8e-4


@pytest.mark.gpu
def test_gpu_rnetwork_forward_tensor_core_mode_matches_reference_golden():
    """R_NETWORK.forward with compute_mode = 'fp16' (the module-level drop-in call on the tensor-core plan) vs the reference."""
    g = torch.load(GOLDEN)
    net = product_net()
    randomise_bn(net.state_dict(), g["bn_seed"])
    net = net.cuda().eval()
    net.compute_mode = "fp16"
    mask = net(torch.abs(O.stft(g["noisy_audio"])).cuda())
    torch.cuda.synchronize()
    assert tuple(mask.shape) == tuple(g["mask"].shape)
    assert rel_err(mask.cpu(), g["mask"]) <= 2e-3
