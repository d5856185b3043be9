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
