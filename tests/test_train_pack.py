"""CPU: the (index, sign) tables of dcsnet_b200.train_pack (the per-step GPU re-packing of the training operands) reproduce
packing.PackedConv / train_ops.dgrad_conv — the same layouts the forward kernels are verified with."""
import pytest
import torch

import dcsnet_b200  # noqa: F401
from dcsnet_b200 import packing, train_ops as T, train_pack as TP


def _leafs(shapes, g):
    vals, syms, off = [], [], 0
    for s in shapes:
        v = torch.randn(*s, generator=g)
        vals.append(v)
        syms.append(TP.Sym.leaf(off, s))
        off += v.numel()
    return vals, syms, torch.cat([v.reshape(-1) for v in vals])


@pytest.mark.parametrize("cin,cout,k,transposed,up", [(1, 8, 7, False, (1, 1)), (16, 32, 5, False, (1, 1)), (64, 64, 3, False, (1, 1)),
                                                     (16, 1, 3, True, (2, 2)), (32, 8, 3, True, (2, 2)), (64, 32, 3, True, (2, 1))])
def test_symbolic_conv_operands_equal_packed_conv(cin, cout, k, transposed, up):
    g = torch.Generator().manual_seed(cin + cout + k)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    (w_r, w_i, b_r, b_i), syms, flat = _leafs([wshape, wshape, (cout,), (cout,)], g)
    pk = packing.PackedConv(w_r, w_i, b_r, b_i, transposed=transposed, up=up, tc_dtype=torch.float16)
    sy = TP.sym_conv(*syms, transposed=transposed, up=up, tc=True)
    assert tuple(sy["w_ffma"].shape) == tuple(pk.w_ffma.shape) and tuple(sy["w_tc"].shape) == tuple(pk.w_tc.shape)
    assert torch.allclose(sy["w_ffma"].evaluate(flat).float(), pk.w_ffma, rtol=1e-6, atol=1e-7)
    assert torch.allclose(sy["bias"].evaluate(flat).float(), pk.bias, rtol=1e-6, atol=1e-7)
    assert torch.equal(sy["w_tc"].evaluate(flat).half(), pk.w_tc)
    idx, sgn = sy["w_ffma"].tables()
    assert idx.shape[1] == 4 and int((idx >= 0).sum(1).max()) <= up[0] * up[1]


@pytest.mark.parametrize("cin,cout,k,transposed", [(8, 16, 7, False), (64, 128, 3, False), (32, 8, 3, True), (128, 128, 1, False)])
def test_symbolic_dgrad_operands_equal_dgrad_conv(cin, cout, k, transposed):
    g = torch.Generator().manual_seed(cin * 3 + cout)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    (w_r, w_i), syms, flat = _leafs([wshape, wshape], g)
    pk = T.dgrad_conv(w_r, w_i, transposed=transposed, device="cpu")
    sy = TP.sym_dgrad(*syms, transposed=transposed)
    assert torch.allclose(sy["w_ffma"].evaluate(flat).float(), pk.w_ffma, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("cin,cout,k,stride", [(8, 16, 7, (2, 2)), (16, 32, 5, (2, 2)), (64, 128, 3, (2, 1))])
def test_symbolic_strided_dgrad_operands_equal_phase_pack(cin, cout, k, stride):
    g = torch.Generator().manual_seed(cin + cout)
    (w_r, w_i), syms, flat = _leafs([(cout, cin, k, k), (cout, cin, k, k)], g)
    pk = T.PhasePack(w_r, w_i, stride, "cpu", want_tf32=True)
    sy = TP.sym_dgrad_strided(*syms, stride=stride, tf32=True)
    assert tuple(sy["w_ffma"].shape) == tuple(pk.w_ffma.shape) and tuple(sy["w_tc32"].shape) == tuple(pk.w_tc32.shape)
    assert torch.allclose(sy["w_ffma"].evaluate(flat).float(), pk.w_ffma, rtol=1e-6, atol=1e-7)
    assert torch.allclose(packing.round_tf32(sy["w_tc32"].evaluate(flat).float()), pk.w_tc32, rtol=1e-6, atol=1e-7)
    # the phase form computes the same data gradient as autograd (a CPU evaluation of the operand semantics)
    x = torch.randn(1, 2 * cin, 8, 12, generator=g, dtype=torch.float64)
    Wp = torch.cat([torch.cat([w_r, -w_i], 1), torch.cat([w_i, w_r], 1)], 0).double()
    OH, OW = (8 + 2 * (k // 2) - k) // stride[0] + 1, (12 + 2 * (k // 2) - k) // stride[1] + 1
    dY = torch.randn(1, 2 * cout, OH, OW, generator=g, dtype=torch.float64)
    want = torch.nn.grad.conv2d_input(x.shape, Wp, dY, stride=stride, padding=k // 2)       # channels [re block | im block]
    got = torch.zeros(1, 8, 12, cin, 2, dtype=torch.float64)
    dyc = torch.stack([dY[0, :cout], dY[0, cout:]], -1).permute(1, 2, 0, 3)                 # (OH, OW, cout, 2)
    W4 = pk.w_ffma.double()                                                                # [p][t][k = (co, ri)][n_pad]
    for ph in range(stride[0]):
        for pw in range(stride[1]):
            p = ph * stride[1] + pw
            for t in range(pk.ntaps):
                d_y, d_x = pk.dy[p * pk.ntaps + t], pk.dx[p * pk.ntaps + t]
                for y in range(8 // stride[0]):
                    for xx in range(12 // stride[1]):
                        if 0 <= y + d_y < OH and 0 <= xx + d_x < OW:
                            v = dyc[y + d_y, xx + d_x].reshape(-1) @ W4[p, t][:, :2 * cin]
                            got[0, y * stride[0] + ph, xx * stride[1] + pw] += v.reshape(cin, 2)
    want_cl = torch.stack([want[0, :cin], want[0, cin:]], -1).permute(1, 2, 0, 3)
    assert torch.allclose(got[0], want_cl, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_whole_model_gather_tables_equal_the_host_packing(mode):
    """TrainStep.init_optimizer on the CPU (flat parameter / gradient buffers, parameters re-pointed as views) builds the (index, sign)
    tables of EVERY operand the training step uses — 14 convs and their data-gradient forms, fc, the stacked LSTM matrices / summed biases,
    the attention gate convs — and each table evaluated on the flat parameters equals what the host packing produced."""
    from dcsnet_b200 import c_network, config as cfg, train_engine
    net = c_network.C_NETWORK(cfg.config, dict(cfg.hparams), 0)
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    step = train_engine.TrainStep(net, "dcs", mode=mode).init_optimizer()
    assert all(torch.equal(before[k], v) for k, v in net.state_dict().items())          # re-pointing the parameters kept their values
    assert int(step.flat_param.numel()) == 2912707 and all(p.data_ptr() >= step.flat_param.data_ptr() for p in net.parameters())
    n = step.gather_tables.check_against_host_packing()
    assert n == {"fp32": 75, "tf32": 112, "bf16": 112}[mode], n
