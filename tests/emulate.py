"""Pure-torch (CPU) emulation of the kernel-side DEFINITIONS in include/dcsnet.h, used to test the host-side
packing without a GPU.  It is test code: it consumes the packed operands exactly as the kernels index them."""
import torch


def conv_geometry(pk, src0, src1, out_hw, weights="ffma"):
    x = src0 if src1 is None else torch.cat([src0, src1], dim=3)
    B, H, W, C, _ = x.shape
    xr = x.reshape(B, H, W, 2 * C).double()
    OH, OW = out_hw
    uh, uw = pk.up
    PH, PW = OH // uh, OW // uw
    out = torch.zeros(B, OH, OW, 2 * pk.cout, dtype=torch.float64)
    if weights == "ffma":
        Wf = pk.w_ffma.double()                                   # [p][t][k][n_pad]
    else:                                                         # tcgen05 operand [p][n_pad][k_pad]
        K = pk.ntaps * 2 * pk.cin
        Wf = pk.w_tc.double()[:, :, :K].reshape(pk.phases, pk.n_pad, pk.ntaps, 2 * pk.cin).permute(0, 2, 3, 1)
    for ph in range(uh):
        for pw in range(uw):
            p = ph * uw + pw
            acc = torch.zeros(B, PH, PW, pk.n_pad, dtype=torch.float64)
            for t in range(pk.ntaps):
                dy, dx = pk.dy[p * pk.ntaps + t], pk.dx[p * pk.ntaps + t]
                jj = torch.arange(PH) * pk.stride[0] + dy
                ii = torch.arange(PW) * pk.stride[1] + dx
                vj, vi = (jj >= 0) & (jj < H), (ii >= 0) & (ii < W)
                g = xr[:, jj.clamp(0, H - 1)][:, :, ii.clamp(0, W - 1)]
                g = g * vj[None, :, None, None] * vi[None, None, :, None]
                acc += g @ Wf[p, t]
            acc = acc + pk.bias.double()
            if pk.act == 1:
                acc = acc.clamp_min(0)
            elif pk.act == 2:
                acc = torch.where(acc > 0, acc, 0.01 * acc)
            out[:, ph::uh, pw::uw] = acc[..., :2 * pk.cout]
    return out.reshape(B, OH, OW, pk.cout, 2)


def nchw_to_cl(t):
    return torch.view_as_real(t.permute(0, 2, 3, 1).contiguous())


def cl_to_nchw(t):
    return torch.view_as_complex(t.float().contiguous()).permute(0, 3, 1, 2)


def lstm_dataflow(lw, x, hidden=64):
    """dcs_clstm_fwd's data flow (lstm.cu) in torch: deinterleave, GEMM, recurrence per (q, dir), layer 1, combine."""
    B, S, D = x.shape
    H = hidden
    rows = B * S
    xp = torch.stack([x.real.reshape(rows, D), x.imag.reshape(rows, D)], 0).reshape(2 * rows, D)
    pre = xp @ lw["w_ih0"] + lw["bias"][:1024]

    def rec(pre_fn, whh):
        hout = torch.zeros(4 * B, S, 2 * H)
        for q in range(4 * B):
            lstm, pb = q // (2 * B), q % (2 * B)
            for d in range(2):
                w = whh[lstm, d]
                h, c = torch.zeros(H), torch.zeros(H)
                for step in range(S):
                    t = S - 1 - step if d else step
                    i, f, g, o = (pre_fn(lstm, pb, t, d) + w @ h).split(H)
                    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                    h = torch.sigmoid(o) * torch.tanh(c)
                    hout[q, t, d * H:(d + 1) * H] = h
        return hout

    h0 = rec(lambda l, pb, t, d: pre[pb * S + t, l * 512 + d * 256:l * 512 + d * 256 + 256], lw["w_hh"][0])
    pre1 = [h0[l * 2 * B:(l + 1) * 2 * B].reshape(2 * rows, 2 * H) @ lw["w_ih1"][l] + lw["bias"][1024 + l * 512:1024 + (l + 1) * 512]
            for l in range(2)]
    h1 = rec(lambda l, pb, t, d: pre1[l][pb * S + t, d * 256:(d + 1) * 256], lw["w_hh"][1])
    n = rows * 2 * H
    f = h1.reshape(-1)
    return torch.complex(f[:n] - f[3 * n:4 * n], f[n:2 * n] + f[2 * n:3 * n]).reshape(B, S, 2 * H)
