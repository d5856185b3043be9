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
            elif pk.act == 3:
                acc = torch.sigmoid(acc)
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


def strip_geometry(sp, src0, src1, out_hw):
    """dcs_cconv2d_strip_fwd's data flow (csrc/cconv_strip.cu) in torch: ring rows, row-shifted / K-sliced strip views,
    swizzled weight image, accumulator columns (ph, pw, n), epilogue scatter.  Consumes sp.item_table / sp.w_image
    exactly as the kernel does (bf16 weights, fp64 accumulation here)."""
    pk = sp.pk
    uh, uw = sp.up
    sh, sw = sp.stride
    N = 2 * pk.cout
    srcs = [src0] + ([src1] if src1 is not None else [])
    B, H, W = src0.shape[:3]
    assert W % sw == 0
    Pb = [4 * sp.c0, 4 * sp.c1]
    P = [b * sw for b in Pb]
    strip_off = [0, (sp.box_units * P[0] + 1023) // 1024 * 1024]
    # strips as the TMA box delivers them: (B, H, W/sw, P/2 bf16 elements)
    rows = [s.reshape(B, H, W // sw, -1).double() for s in srcs]
    OH, OW = out_hw
    PH, PW = OH // uh, OW // uw
    out = torch.zeros(B, OH, OW, N, dtype=torch.float64)
    img = sp.w_image.cpu()
    tab = sp.item_table.cpu().to(torch.int64)
    n_strips = (PW + 127) // 128
    run = uw * N
    for g in sp.groups:
        items = tab[g["item0"]:g["item0"] + g["n_items"]]
        for strip in range(n_strips):
            x0 = strip * 128
            for j in range(PH):
                acc = torch.full((B, 128, sp.cols), float("nan"), dtype=torch.float64)
                for a16, b16, z, _ in items.tolist():
                    d_col, drow, flags = z & 0xffff, (z >> 16) & 0xff, (z >> 24) & 0xff
                    s = 1 if flags & 2 else 0
                    rel = a16 * 16 - strip_off[s]
                    shift, within = rel // P[s], rel % P[s]
                    assert 0 <= shift and shift + 128 <= sp.box_units and within % 32 == 0
                    y = j * sh + g["dy_min"] + drow
                    ux = x0 + g["x_min"] + shift + torch.arange(128)
                    ok = (ux >= 0) & (ux < W // sw) & (0 <= y < H)
                    A = rows[s][:, min(max(y, 0), H - 1)][:, ux.clamp(0, W // sw - 1)][..., within // 2:within // 2 + 16]
                    A = A * ok[None, :, None]
                    n = torch.arange(sp.n_mma)[:, None]
                    k = torch.arange(16)[None, :]
                    off = n * 32 + k * 2
                    off = off ^ (((off >> 7) & 1) << 4)
                    Wb = img[(g["w_off"] + b16 * 16 + off) // 2].double()          # [n_mma][16]
                    part = A @ Wb.t()
                    if flags & 1:
                        acc[:, :, d_col:d_col + sp.n_mma] = part
                    else:
                        acc[:, :, d_col:d_col + sp.n_mma] += part
                acc = acc + pk.bias.double()[torch.arange(sp.cols) % N]
                if pk.act == 1:
                    acc = acc.clamp_min(0)
                elif pk.act == 2:
                    acc = torch.where(acc > 0, acc, 0.01 * acc)
                nx = min(128, PW - x0)
                for c0 in range(0, sp.cols, run):
                    oy = j * uh + g["ph0"] + c0 // run
                    out[:, oy, x0 * uw:(x0 + nx) * uw] = acc[:, :nx, c0:c0 + run].reshape(B, nx * uw, N)
    return out.reshape(B, OH, OW, pk.cout, 2)


def rnet_dataflow(pk, mag):
    """R_NETWORK.forward (r_network.py:125-173) evaluated through packing.PackedRNet: every conv / fc / BatchNorm goes
    through the kernel operand definitions (conv_geometry on channel-pair tensors), the real CBAM and the real LSTM use the
    packed weights with plain torch arithmetic (the definitions the round-2 kernels will implement).  mag: (B, F, T) fp32."""
    import torch.nn.functional as F
    B, Fb, T = mag.shape
    Lr = pk.L

    def out_hw(p, H, W):
        if p.up != (1, 1):
            return H * p.up[0], W * p.up[1]
        return (H + 2 * (p.kh // 2) - p.kh) // p.stride[0] + 1, (W + 2 * (p.kw // 2) - p.kw) // p.stride[1] + 1

    def real(t):      # (B,H,W,C/2,2) channel pairs -> (B,H,W,C) real channels
        return t.reshape(t.shape[0], t.shape[1], t.shape[2], -1)

    def attention(x, ca, w7):
        xr = real(x)                                                  # (B,H,W,C)
        m = xr.amax(dim=(1, 2))                                       # global max pool (the only branch that counts)
        gate = torch.sigmoid(F.relu(m @ ca["w1"].double().T) @ ca["w2"].double().T)   # (B,C)
        u = xr * gate[:, None, None, :]
        st = torch.stack([u.mean(dim=3), u.amax(dim=3)], dim=1)      # (B,2,H,W)
        s = torch.sigmoid(F.conv2d(st, w7.double().reshape(1, 2, 7, 7), padding=3))[:, 0]
        return (u * s[..., None]).reshape(x.shape)

    a6 = pk.bn0.double()[0]
    x = torch.zeros(B, Fb, T, 1, 2, dtype=torch.float64)
    x[..., 0, 0] = a6[0] * mag.double() + a6[4]                       # initial BatchNorm2d on the magnitude (pair's .im stays 0)
    enc = [x]
    H, W = Fb, T
    for i in range(Lr):
        H, W = out_hw(pk.enc[i], H, W)
        enc.append(conv_geometry(pk.enc[i], enc[i], None, (H, W)))
    e = real(enc[-1])                                                 # (B,H,W,256)
    seq = e.reshape(B, H * W, -1)
    for layer in range(2):
        outs = []
        for d, w in enumerate(pk.lstm[layer]):
            Hd = w["w_hh"].shape[1]
            h = torch.zeros(B, Hd, dtype=torch.float64)
            c = torch.zeros(B, Hd, dtype=torch.float64)
            out = torch.zeros(B, seq.shape[1], Hd, dtype=torch.float64)
            steps = range(seq.shape[1] - 1, -1, -1) if d else range(seq.shape[1])
            for t in steps:
                g = seq[:, t] @ w["w_ih"].double().T + h @ w["w_hh"].double().T + w["bias"].double()
                i_, f_, g_, o_ = g.chunk(4, dim=1)
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                h = torch.sigmoid(o_) * torch.tanh(c)
                out[:, t] = h
            outs.append(out)
        seq = torch.cat(outs, dim=2)
    lat = seq.reshape(B, 1, H * W, -1, 2)                             # sequence index = h * W + w (flatten(2, 3))
    d = conv_geometry(pk.fc, lat, None, (1, H * W)).reshape(B, H, W, -1, 2)
    for i in range(Lr):
        skip = attention(enc[Lr - i], *pk.skip_att[i])
        H, W = out_hw(pk.dec[i], H, W)
        d = conv_geometry(pk.dec[i], d, skip, (H, W))
        if i != Lr - 1:
            d = attention(d, *pk.dec_att[i])
    return d[..., 0, 0]                                               # 1-channel output in the pair's .re (sigmoid = the packed activation)
