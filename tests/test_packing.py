"""CPU: host-side weight packing (BN fold, bias rule, flipped transposed conv, sub-pixel phases, two-source K order,
LSTM / attention layouts) checked through a torch emulation of the kernel-side definitions against the oracle."""
import pytest
import torch

import dcsnet_b200 as D
from dcsnet_b200 import packing, ops
from oracle import dcsnet_oracle as O, synthetic_weights as SW
from emulate import conv_geometry, strip_geometry, nchw_to_cl, cl_to_nchw, lstm_dataflow
from conftest import rel_err


@pytest.fixture(scope="module")
def packed():
    sd = SW.make_state_dict(0)
    return sd, D.PackedNet(sd, "cpu", "bf16")


def _rand_c(g, *shape):
    return torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g))


@pytest.mark.parametrize("i", range(7))
def test_encoder_layer_packing(packed, i):
    sd, pk = packed
    g = torch.Generator().manual_seed(i)
    p = pk.enc[i]
    x = _rand_c(g, 2, p.cin, 16, 24)
    ref = O.crelu(O.cbn_eval(O.cconv2d(x, sd, f"encoder.{i}.0.", O.STRIDE_E[i], O.KERNEL_E[i] // 2), sd, f"encoder.{i}.1."))
    got = cl_to_nchw(conv_geometry(p, nchw_to_cl(x), None, ops.conv_out_hw(p, 16, 24)))
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= 5e-6


@pytest.mark.parametrize("i", range(7))
def test_decoder_layer_packing(packed, i):
    """cat(d, skip) -> nearest upsample -> ConvTranspose2d(k3,s1,p1) [-> BN -> LReLU]  ==  phase-decomposed two-source GEMM."""
    sd, pk = packed
    g = torch.Generator().manual_seed(10 + i)
    p = pk.dec[i]
    c = p.cin // 2
    d, s = _rand_c(g, 2, c, 4, 6), _rand_c(g, 2, c, 4, 6)
    u = O.cupsample_nearest(torch.cat((d, s), 1), O.UPSAMPLE[i])
    if i == 6:
        ref = O.cconvT2d(u, sd, "decoder.6.", 1, 1)
    else:
        ref = O.clrelu(O.cbn_eval(O.cconvT2d(u, sd, f"decoder.{i}.0.", 1, 1), sd, f"decoder.{i}.1."))
    got = cl_to_nchw(conv_geometry(p, nchw_to_cl(d), nchw_to_cl(s), ref.shape[2:]))
    assert rel_err(got, ref) <= 5e-6
    # executed MACs drop by 1.5x / 2.25x with the pre-summed taps
    assert p.phases * p.ntaps == {(2, 1): 12, (2, 2): 16}[p.up]


def test_tc_operand_is_rounded_ffma_operand(packed):
    _, pk = packed
    for p in pk.enc[1:] + pk.dec:
        K = p.ntaps * 2 * p.cin
        assert p.w_tc.shape == (p.phases, p.n_pad, (K + 63) // 64 * 64)
        assert torch.all(p.w_tc[:, :, K:] == 0)
        wt = p.w_tc.float()[:, :, :K].reshape(p.phases, p.n_pad, p.ntaps, 2 * p.cin).permute(0, 2, 3, 1)
        assert torch.equal(wt, p.w_ffma.to(torch.bfloat16).float())


def test_fc_packing(packed):
    sd, pk = packed
    g = torch.Generator().manual_seed(5)
    x = _rand_c(g, 2, 10, 128)
    ref = O.clinear(x, sd, "fc.")
    got = conv_geometry(pk.fc, torch.view_as_real(x).reshape(2, 1, 10, 128, 2), None, (1, 10))
    got = torch.view_as_complex(got.float().contiguous()).reshape(2, 10, 128)
    assert rel_err(got, ref) <= 5e-6


def test_bn_fold_is_eval_bn(packed):
    sd, _ = packed
    g = torch.Generator().manual_seed(6)
    x = _rand_c(g, 2, 16, 5, 7)
    A, c = packing.bn_affine_from_sd(sd, "encoder.1.1.")
    xr = torch.stack([x.real, x.imag], -1).double()                       # (B,C,H,W,2)
    y = torch.einsum("cab,nchwb->nchwa", A, xr) + c[None, :, None, None, :]
    ref = O.cbn_eval(x, sd, "encoder.1.1.")
    assert rel_err(torch.complex(y[..., 0], y[..., 1]).to(torch.complex64), ref) <= 5e-6


def test_lstm_packing_and_dataflow(packed):
    sd, pk = packed
    g = torch.Generator().manual_seed(2)
    x = _rand_c(g, 3, 7, 128)
    ref = O.complex_lstm(x, sd, "lstm.", explicit=True)
    assert rel_err(lstm_dataflow(pk.lstm, x), ref) <= 5e-6


def test_bias_rule_quirk():
    """complexPyTorch's apply_complex gives the effective bias (b_r - b_i) + j (b_r + b_i), not b_r + j b_i."""
    wr, wi = torch.zeros(3, 2, 1, 1), torch.zeros(3, 2, 1, 1)
    br, bi = torch.tensor([1.0, 2.0, 3.0]), torch.tensor([0.5, -1.0, 4.0])
    p = packing.PackedConv(wr, wi, br, bi)
    assert torch.allclose(p.bias[:6].view(3, 2), torch.stack([br - bi, br + bi], 1))


@pytest.mark.parametrize("layer,merged,groups", [("enc1", True, 1), ("enc2", True, 1), ("dec5", True, 1), ("dec5", False, 1),
                                                 ("dec4", False, 2), ("dec4", True, 2)])
def test_strip_packing(packed, layer, merged, groups):
    """The row-strip kernel's item table + swizzled weight image (packing.StripConv) describe the same convolution as
    the per-tap operands (conv_geometry), up to the bf16 rounding of the weights."""
    _, pk = packed
    p = {"enc1": pk.enc[1], "enc2": pk.enc[2], "dec4": pk.dec[4], "dec5": pk.dec[5], "dec6": pk.dec[6]}[layer]
    g = torch.Generator().manual_seed(31)
    if layer in ("enc1", "enc2"):
        H, W = 6, 260                      # W/2 = 130 output columns: two strips, ragged
        srcs = [torch.randn(2, H, W, p.cin, 2, generator=g), None]
        c0, c1 = p.cin, 0
    else:
        H, W = 3, 131
        c = p.cin // 2
        srcs = [torch.randn(2, H, W, c, 2, generator=g), torch.randn(2, H, W, c, 2, generator=g)]
        c0, c1 = c, c
    srcs = [None if s is None else s.to(torch.bfloat16).float() for s in srcs]
    sp = packing.StripConv(p, c0, c1, merged=merged, groups=groups)
    out_hw = ops.conv_out_hw(p, H, W)
    ref = conv_geometry(p, srcs[0], srcs[1], out_hw, weights="tc")
    got = strip_geometry(sp, srcs[0], srcs[1], out_hw)
    assert not torch.isnan(got).any()
    assert rel_err(got, ref) <= 1e-2     # merged blocks sum bf16-rounded pre-summed taps in a different order: bf16-level
    # and against the unrounded operands at bf16 accuracy
    assert rel_err(got, conv_geometry(p, srcs[0], srcs[1], out_hw)) <= 2e-2


def test_strip_enc0_toeplitz_packing(packed):
    """encoder[0] on the row-strip kernel: K taken from space (16-pixel strip rows, Toeplitz weight blocks)."""
    _, pk = packed
    p = pk.enc[0]
    g = torch.Generator().manual_seed(3)
    B, F, T = 2, 12, 48
    x = torch.randn(B, F, T, 1, 2, generator=g).to(torch.bfloat16).float()
    sp = packing.StripEnc0(p)
    out_hw = ops.conv_out_hw(p, F, T)
    ref = conv_geometry(p, x, None, out_hw)
    got = strip_geometry(sp, packing.StripEnc0.view_src(x), None, out_hw)
    assert not torch.isnan(got).any()
    assert rel_err(got, ref) <= 1e-2


def test_strip_dec6_toeplitz_packing(packed):
    """decoder[6] on the row-strip kernel: N taken from space (4-pixel strip rows, Toeplitz blocks of pre-summed taps)."""
    _, pk = packed
    p = pk.dec[6]
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 3, 136                     # W/4 = 34 strip rows
    d = torch.randn(B, H, W, 8, 2, generator=g).to(torch.bfloat16).float()
    k = torch.randn(B, H, W, 8, 2, generator=g).to(torch.bfloat16).float()
    sp = packing.StripDec6(p)
    ref = conv_geometry(p, d, k, (2 * H, 2 * W))
    got = strip_geometry(sp, sp.view_src(d), sp.view_src(k), (2 * H, 2 * W))
    assert not torch.isnan(got).any()
    assert rel_err(got, ref) <= 1e-2
