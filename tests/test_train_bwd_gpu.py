"""GPU (B200): the second slice of the training-step kernels (csrc/train_bwd.cu; include/dcsnet.h "f2, second slice") against the
closed-form contracts of oracle/train_oracle.py (each itself checked against autograd on the CPU, tests/test_train_oracle.py) and
against torch reference ops (torch.nn.LSTM autograd, torch.optim.Adam).  'rel' = max|a-b| / max|b|, fp32 kernels."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dcsnet_oracle as O, train_oracle as TO  # noqa: E402
from conftest import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu


def cl(t):       # NCHW complex -> channels-last (B, H, W, C, 2) fp32
    return torch.view_as_real(t.permute(0, 2, 3, 1).contiguous())


def nchw(t):     # channels-last (B, H, W, C, 2) -> NCHW complex
    return torch.view_as_complex(t.detach().float().cpu().contiguous()).permute(0, 3, 1, 2)


def rc(g, *s):
    return torch.complex(torch.randn(*s, generator=g), torch.randn(*s, generator=g))


@pytest.mark.parametrize("cin,cout,k,stride,H,W", [(1, 8, 7, (2, 2), 32, 40), (8, 16, 7, (2, 2), 16, 36), (16, 32, 5, (2, 2), 12, 22),
                                                  (32, 64, 5, (2, 1), 8, 19), (64, 128, 3, (2, 1), 6, 17), (128, 128, 3, (2, 1), 4, 9)])
def test_encoder_conv_backward_generic_wgrad_bias_and_strided_dgrad(cin, cout, k, stride, H, W):
    """ComplexConv2d(k, stride, p = k // 2) backward: dcs_wgrad + dcs_wgrad_fold_complex (weights), dcs_colsum mode 1 (biases),
    dcs_dilate + the forward FFMA conv with the flipped in / out-swapped weights (data) vs oracle cconv2d_backward."""
    from dcsnet_b200 import ops, train_ops as T
    g = torch.Generator().manual_seed(cin * 7 + cout)
    B = 2
    x = rc(g, B, cin, H, W)
    w_r, w_i = 0.1 * torch.randn(cout, cin, k, k, generator=g), 0.1 * torch.randn(cout, cin, k, k, generator=g)
    OH, OW = (H + 2 * (k // 2) - k) // stride[0] + 1, (W + 2 * (k // 2) - k) // stride[1] + 1
    dy = rc(g, B, cout, OH, OW)
    dx_w, dwr_w, dwi_w, dbr_w, dbi_w = TO.cconv2d_backward(x, w_r, w_i, dy, stride, k // 2)
    xc, dyc = cl(x).cuda(), cl(dy).cuda()
    dwr, dwi = T.cwgrad_generic(xc, dyc, k, stride)
    dbr, dbi = T.colsum(dyc.view(-1, 2 * cout), mode=1)
    pk = T.dgrad_conv(w_r, w_i, transposed=False, device="cuda")
    dx = ops.cconv(pk, T.dilate(dyc, H, W, stride), None, torch.empty(B, H, W, cin, 2, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    assert rel_err(dwr, dwr_w) <= 2e-5 and rel_err(dwi, dwi_w) <= 2e-5
    assert rel_err(dbr, dbr_w) <= 2e-5 and rel_err(dbi, dbi_w) <= 2e-5
    assert rel_err(nchw(dx), dx_w) <= 2e-5
    if (H, W) == (OH * stride[0], OW * stride[1]):      # the phase form (no zero insertion): FFMA and tcgen05 kind::tf32
        pp = T.PhasePack(w_r, w_i, stride, "cuda", want_tf32=True)
        dxp = ops.cconv(pp, dyc, None, torch.empty(B, H, W, cin, 2, dtype=torch.float32, device="cuda"))
        assert rel_err(nchw(dxp), dx_w) <= 2e-5
        if cout % 4 == 0 and 2 * cin <= 256:
            dxt = ops.cconv(pp, dyc, None, torch.empty(B, H, W, cin, 2, dtype=torch.float32, device="cuda"), use_tc=True)
            assert rel_err(nchw(dxt), dx_w) <= 2e-3
    if cin == 1:      # encoder[0]: the direct gather kernel on the raw weights
        dx1 = T.cconv_dgrad_cin1(dyc, w_r.cuda(), w_i.cuda(), H, W, stride)
        assert rel_err(nchw(dx1), dx_w) <= 2e-5


@pytest.mark.parametrize("cd,cs,cout,up", [(8, 8, 1, (2, 2)), (16, 16, 8, (2, 2)), (64, 64, 32, (2, 1)), (128, 128, 128, (2, 1))])
def test_decoder_stage_weight_and_bias_gradients(cd, cs, cout, up):
    """cat + nearest up-sampling + ComplexConvTranspose2d(k3 s1 p1): dcs_upcat_fwd materialises the conv's input, dcs_wgrad folds with
    transposed = 1 into the module's (in, out, 3, 3) layout; vs oracle decoder_stage_backward."""
    from dcsnet_b200 import train_ops as T
    g = torch.Generator().manual_seed(cd + cout)
    B, H, W = 2, 5, 11
    d, skip = rc(g, B, cd, H, W), rc(g, B, cs, H, W)
    w_r, w_i = 0.1 * torch.randn(cd + cs, cout, 3, 3, generator=g), 0.1 * torch.randn(cd + cs, cout, 3, 3, generator=g)
    dy = rc(g, B, cout, H * up[0], W * up[1])
    _, _, dwr_w, dwi_w, dbr_w, dbi_w = TO.decoder_stage_backward(d, skip, w_r, w_i, dy, up)
    z = T.upcat_fwd(cl(d).cuda(), cl(skip).cuda(), up)
    assert rel_err(nchw(z), O.cupsample_nearest(torch.cat([d, skip], 1), up)) == 0
    dyc = cl(dy).cuda()
    dwr, dwi = T.cwgrad_generic(z, dyc, 3, (1, 1), transposed=True)
    dbr, dbi = T.colsum(dyc.view(-1, 2 * cout), mode=1)
    torch.cuda.synchronize()
    assert rel_err(dwr, dwr_w) <= 2e-5 and rel_err(dwi, dwi_w) <= 2e-5
    assert rel_err(dbr, dbr_w) <= 2e-5 and rel_err(dbi, dbi_w) <= 2e-5


def test_act_bwd_and_dropout_kernels():
    from dcsnet_b200 import _lib as L, train_ops as T
    g = torch.Generator().manual_seed(11)
    y, g0, g1 = (torch.randn(2, 5, 7, 8, 2, generator=g).cuda() for _ in range(3))
    cc = torch.randn(2, 8, 2, generator=g).cuda()
    tot = g0 + g1 + cc[:, None, None]
    assert torch.equal(T.act_bwd(y, g0, L.ACT_RELU, g1, cc), tot * (y > 0))
    assert torch.allclose(T.act_bwd(y, g0, L.ACT_LRELU, g1, cc), tot * torch.where(y > 0, 1.0, 0.01), rtol=1e-6, atol=0)
    assert torch.equal(T.act_bwd(None, g0, L.ACT_NONE), g0)
    x = torch.randn(1 << 20, generator=g).cuda()
    a, b = T.dropout(x, 0.1, seed=5, offset=0), T.dropout(x, 0.1, seed=5, offset=0)
    assert torch.equal(a, b)                                                     # same (seed, offset): same mask (the backward's contract)
    keep = a != 0
    assert abs(float(keep.float().mean()) - 0.9) < 2e-3
    assert torch.allclose(a[keep], x[keep] / 0.9, rtol=1e-6)
    assert not torch.equal(T.dropout(x, 0.1, seed=5, offset=1 << 20) != 0, keep) and not torch.equal(T.dropout(x, 0.1, seed=6, offset=0) != 0, keep)
    assert torch.equal(T.dropout(x, 0.0, seed=1, offset=0), x)
    pairs = keep.view(-1, 2)                                                     # real and imaginary parts drop independently
    assert abs(float((pairs[:, 0] & pairs[:, 1]).float().mean()) - 0.81) < 3e-3


@pytest.mark.parametrize("C,H,W", [(8, 12, 20), (16, 9, 33), (32, 8, 10), (128, 4, 9)])
def test_attention_backward_kernels(C, H, W):
    """dcs_attention_bwd (+ the per-channel constant) vs oracle attention_backward: dx and all six weight gradients."""
    from dcsnet_b200 import packing, train_ops as T
    g = torch.Generator().manual_seed(C + H)
    B, R = 2, max(C // 16, 1)
    x, dy = rc(g, B, C, H, W), rc(g, B, C, H, W)
    sd = {"c.fc.0.conv_r.weight": 0.5 * torch.randn(R, C, 1, 1, generator=g), "c.fc.0.conv_i.weight": 0.5 * torch.randn(R, C, 1, 1, generator=g),
          "c.fc.2.conv_r.weight": 0.5 * torch.randn(C, R, 1, 1, generator=g), "c.fc.2.conv_i.weight": 0.5 * torch.randn(C, R, 1, 1, generator=g),
          "s.conv1.conv_r.weight": 0.2 * torch.randn(1, 2, 7, 7, generator=g), "s.conv1.conv_i.weight": 0.2 * torch.randn(1, 2, 7, 7, generator=g)}
    dx_w, gw = TO.attention_backward(x, sd, "c.", "s.", dy)
    ca, w7 = packing.pack_channel_attention(sd, "c.", "cuda"), packing.pack_spatial_attention(sd, "s.", "cuda")
    y, sv = T.attention_fwd_saved(cl(x).cuda(), ca, w7)
    dx, cc, gr = T.attention_bwd(sv["x"], cl(dy).cuda(), sv["gate_c"], sv["stats"], sv["gate_s"], sv["sums"], ca, w7)
    dx = T.act_bwd(None, dx, 0, None, cc)
    torch.cuda.synchronize()
    assert rel_err(nchw(dx), dx_w) <= 5e-5
    for ours, name in (("dw1_r", "c.fc.0.conv_r.weight"), ("dw1_i", "c.fc.0.conv_i.weight"), ("dw2_r", "c.fc.2.conv_r.weight"),
                       ("dw2_i", "c.fc.2.conv_i.weight"), ("dw7_r", "s.conv1.conv_r.weight"), ("dw7_i", "s.conv1.conv_i.weight")):
        assert rel_err(gr[ours].reshape(gw[name].shape), gw[name]) <= 5e-5, name


def test_sgemm_and_transpose():
    from dcsnet_b200 import train_ops as T
    g = torch.Generator().manual_seed(2)
    A, Bn, bias = torch.randn(150, 70, generator=g), torch.randn(90, 70, generator=g), torch.randn(90, generator=g)
    out = T.sgemm(A.cuda(), Bn.cuda(), bias.cuda())
    assert rel_err(out, A @ Bn.t() + bias) <= 1e-5
    Bk = torch.randn(70, 33, generator=g)
    out2 = T.sgemm(A.cuda(), Bk.cuda(), b_is_nk=False)
    assert rel_err(out2, A @ Bk) <= 1e-5
    T.sgemm(A.cuda(), Bk.cuda(), out=out2, b_is_nk=False, accumulate=True)
    assert rel_err(out2, 2 * (A @ Bk)) <= 1e-5
    wide = torch.randn(37, 100, generator=g).cuda()
    dst = torch.empty(45, 37, device="cuda")
    T.transpose_into(wide[:, 20:65], dst)
    assert torch.equal(dst, wide[:, 20:65].t())


@pytest.mark.parametrize("S", [7, 50])
def test_lstm_training_forward_and_bptt(S):
    """dcs_lstm_train_fwd / dcs_lstm_train_bwd + the GEMM-side gradients (dW_hh through dcs_wgrad with a one-step tap) vs the oracle's
    lstm_forward_saved / lstm_bptt (both directions, two weight groups)."""
    from dcsnet_b200 import train_ops as T
    g = torch.Generator().manual_seed(S)
    Q, H, groups = 6, 64, 2
    pre = torch.randn(Q, S, 2, 4 * H, generator=g)
    whh = 0.2 * torch.randn(groups, 2, 4 * H, H, generator=g)
    dh = torch.randn(Q, S, 2, H, generator=g)
    h, gates, cells = T.lstm_train_fwd(pre.cuda(), whh.cuda(), groups)
    dpre = T.lstm_train_bwd(whh.cuda(), gates, cells, dh.cuda(), groups)
    torch.cuda.synchronize()
    for grp in range(groups):
        qs = slice(grp * Q // groups, (grp + 1) * Q // groups)
        for d in range(2):
            flip = (lambda t: t.flip(1)) if d else (lambda t: t)
            hs_w, cs_w, gates_w = TO.lstm_forward_saved(flip(pre[qs, :, d]), whh[grp, d])
            assert rel_err(flip(h[qs, :, d].cpu()), hs_w) <= 2e-5 and rel_err(flip(gates[qs, :, d].cpu()), gates_w) <= 2e-5
            dpre_w, dW_w = TO.lstm_bptt(whh[grp, d], hs_w, cs_w, gates_w, flip(dh[qs, :, d]))
            assert rel_err(flip(dpre[qs, :, d].cpu()), dpre_w) <= 5e-5
            # dW_hh = sum_t da_t^T h_{t-1}: one tap shifted by one step along the sequence
            nq = Q // groups
            dwp = T.wgrad(h[qs].view(nq, 1, S, 2 * H)[..., d * H:(d + 1) * H], dpre[qs].view(nq, 1, S, 8 * H)[..., d * 4 * H:(d + 1) * 4 * H],
                          [(0, 1 if d else -1)])
            assert rel_err(dwp[0].t(), dW_w) <= 5e-5


def test_clstm_split_merge_combine():
    from dcsnet_b200 import train_ops as T
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 10, 16, 2, generator=g).cuda()
    pl = T.cplx_split(x)
    assert torch.equal(pl[0], x[..., 0]) and torch.equal(pl[1], x[..., 1]) and torch.equal(T.cplx_merge(pl), x)
    h = torch.randn(2, 2, 3, 10, 16, generator=g).cuda()          # (lstm R / I, part re / im, ...)
    out = T.clstm_combine(h)
    assert torch.equal(out[..., 0], h[0, 0] - h[1, 1]) and torch.equal(out[..., 1], h[0, 1] + h[1, 0])
    dh = T.clstm_combine_bwd(x)
    assert torch.equal(dh[0, 0], x[..., 0]) and torch.equal(dh[1, 1], -x[..., 0]) and torch.equal(dh[0, 1], x[..., 1]) and torch.equal(dh[1, 0], x[..., 1])


@pytest.mark.parametrize("cin,cout,k,stride,H,W", [(8, 16, 7, (2, 2), 16, 140), (16, 32, 5, (2, 2), 12, 70), (32, 8, 3, (1, 1), 6, 130), (64, 16, 3, (1, 1), 5, 66),
                                                  (256, 128, 3, (1, 1), 4, 40), (256, 64, 3, (1, 1), 4, 33), (128, 128, 1, (1, 1), 1, 200)])
def test_generic_interface_wgrad_on_tcgen05(cin, cout, k, stride, H, W):
    """dcs_wgrad_tc16 (the tcgen05 K = pixels GEMM behind dcs_wgrad's interface: TMA zero-fill pads few-channel operands to the 128 x 64
    tile, > 256 real input channels run as channel slices) against dcs_wgrad on the same bf16-rounded operands."""
    from dcsnet_b200 import train_ops as T
    g = torch.Generator().manual_seed(cin + 3 * cout + k)
    B = 2
    x = torch.randn(B, H, W, 2 * cin, generator=g).bfloat16()
    OH, OW = (H + 2 * (k // 2) - k) // stride[0] + 1, (W + 2 * (k // 2) - k) // stride[1] + 1
    dy = (0.1 * torch.randn(B, OH, OW, 2 * cout, generator=g)).bfloat16()
    taps = T.conv_taps(k, k)
    want = T.wgrad(x.float().cuda(), dy.float().cuda(), taps, stride)
    got = T.wgrad_tc16(x.cuda(), dy.cuda(), taps, stride)
    torch.cuda.synchronize()
    assert rel_err(got, want) <= 2e-5


@pytest.mark.parametrize("H,W", [(8, 32), (5, 41), (16, 70)])
def test_fused_decoder6_backward(H, W):
    """dcs_dec6_bwd (data + weight + bias gradients of decoder[6] in one kernel) vs oracle decoder_stage_backward (cout 1, up (2,2))."""
    from dcsnet_b200 import train_ops as T
    g = torch.Generator().manual_seed(H * W)
    B = 3
    d, skip = rc(g, B, 8, H, W), rc(g, B, 8, H, W)
    w_r, w_i = 0.3 * torch.randn(16, 1, 3, 3, generator=g), 0.3 * torch.randn(16, 1, 3, 3, generator=g)
    dy = rc(g, B, 1, 2 * H, 2 * W)
    gd_w, gs_w, dwr_w, dwi_w, dbr_w, dbi_w = TO.decoder_stage_backward(d, skip, w_r, w_i, dy, (2, 2))
    dwr, dwi = torch.empty(16, 1, 3, 3, device="cuda"), torch.empty(16, 1, 3, 3, device="cuda")
    dbr, dbi = torch.empty(1, device="cuda"), torch.empty(1, device="cuda")
    gd, gs = T.dec6_bwd(cl(d).cuda(), cl(skip).cuda(), cl(dy).cuda(), w_r.cuda(), w_i.cuda(), dwr, dwi, dbr, dbi)
    torch.cuda.synchronize()
    assert rel_err(nchw(gd), gd_w) <= 2e-5 and rel_err(nchw(gs), gs_w) <= 2e-5
    assert rel_err(dwr, dwr_w) <= 2e-5 and rel_err(dwi, dwi_w) <= 2e-5
    assert rel_err(dbr, dbr_w) <= 2e-5 and rel_err(dbi, dbi_w) <= 2e-5


def test_backward_kernels_read_saved_activations_in_16_bit_storage():
    """The bf16 training mode stores the saved forward activations in bf16; the backward kernels that read them (attention backward,
    BN backward, activation mask, fused decoder[6] backward, up-sampled concat, complex split, dropout) must give what they give on the
    same values held in fp32."""
    from dcsnet_b200 import _lib as L, packing, train_ops as T
    g = torch.Generator().manual_seed(21)
    bf = lambda t: t.bfloat16()                                        # noqa: E731
    # attention backward, C = 8 (thread-per-pixel kernels) and C = 32 (lane-per-channel kernels)
    for C in (8, 32):
        B, H, W, R = 2, 6, 20, max(C // 16, 1)
        x16 = bf(torch.randn(B, H, W, C, 2, generator=g)).cuda()
        dy = torch.randn(B, H, W, C, 2, generator=g).cuda()
        sd = {"c.fc.0.conv_r.weight": 0.5 * torch.randn(R, C, 1, 1, generator=g), "c.fc.0.conv_i.weight": 0.5 * torch.randn(R, C, 1, 1, generator=g),
              "c.fc.2.conv_r.weight": 0.5 * torch.randn(C, R, 1, 1, generator=g), "c.fc.2.conv_i.weight": 0.5 * torch.randn(C, R, 1, 1, generator=g),
              "s.conv1.conv_r.weight": 0.2 * torch.randn(1, 2, 7, 7, generator=g), "s.conv1.conv_i.weight": 0.2 * torch.randn(1, 2, 7, 7, generator=g)}
        ca, w7 = packing.pack_channel_attention(sd, "c.", "cuda"), packing.pack_spatial_attention(sd, "s.", "cuda")
        _, sv = T.attention_fwd_saved(x16.float(), ca, w7)             # gates / statistics from the fp32 copy of the same values
        a = T.attention_bwd(x16.float(), dy, sv["gate_c"], sv["stats"], sv["gate_s"], sv["sums"], ca, w7)
        b = T.attention_bwd(x16, dy, sv["gate_c"], sv["stats"], sv["gate_s"], sv["sums"], ca, w7)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and all(torch.equal(a[2][k], b[2][k]) for k in a[2])
    # BN backward, activation mask
    x16 = bf(torch.randn(2, 5, 9, 16, 2, generator=g)).cuda()
    dy = torch.randn(2, 5, 9, 16, 2, generator=g).cuda()
    w, bb = (torch.rand(16, 3, generator=g) + torch.tensor([0.5, 0.5, -0.5])).cuda(), torch.randn(16, 2, generator=g).cuda()
    _, saved, _ = T.cbn_train_fwd(x16.float(), w, bb)
    ra, rb = T.cbn_train_bwd(x16.float(), dy, saved, w), T.cbn_train_bwd(x16, dy, saved, w)
    assert all(torch.equal(p, q) for p, q in zip(ra, rb))
    assert torch.equal(T.act_bwd(x16, dy, L.ACT_LRELU), T.act_bwd(x16.float(), dy, L.ACT_LRELU))
    # fused decoder[6] backward, up-sampled concat, complex split, dropout
    d16, s16 = bf(torch.randn(2, 6, 20, 8, 2, generator=g)).cuda(), bf(torch.randn(2, 6, 20, 8, 2, generator=g)).cuda()
    dpre = torch.randn(2, 12, 40, 1, 2, generator=g).cuda()
    w_r, w_i = torch.randn(16, 1, 3, 3, generator=g).cuda(), torch.randn(16, 1, 3, 3, generator=g).cuda()
    outs = []
    for d_, s_ in ((d16, s16), (d16.float(), s16.float())):
        gr = [torch.empty(16, 1, 3, 3, device="cuda"), torch.empty(16, 1, 3, 3, device="cuda"), torch.empty(1, device="cuda"), torch.empty(1, device="cuda")]
        outs.append(list(T.dec6_bwd(d_, s_, dpre, w_r, w_i, *gr)) + gr)
    assert all(torch.equal(p, q) for p, q in zip(*outs))
    assert torch.equal(T.upcat_fwd(d16, s16, (2, 2), dtype=torch.bfloat16), T.upcat_fwd(d16.float(), s16.float(), (2, 2), dtype=torch.bfloat16))
    assert torch.equal(T.cplx_split(d16), T.cplx_split(d16.float()))
    m16, m32 = T.dropout(d16, 0.1, 3, 17), T.dropout(d16.float(), 0.1, 3, 17)
    assert torch.equal(m16 != 0, m32 != 0) and torch.equal(m16, m32.bfloat16())
