"""Host-side weight packing: reference-format parameters -> GEMM-ready operands.

Done once per (state_dict, mode) in float64 on the host, then moved to the device.  Every fold is exact algebra:
  * complex conv -> real block matrix [[Wr, -Wi], [Wi, Wr]] over interleaved (re, im) channels, with the
    complexPyTorch bias rule (b_r - b_i, b_r + b_i)                               (apply_complex, Appendix A1)
  * ConvTranspose2d(k3, s1, p1) -> flipped, in/out-swapped convolution            (c_network.py:135-147)
  * eval-mode ComplexBatchNorm2d -> per-channel 2x2 affine folded into W and bias (Appendix A3)
  * nearest up-sampling in front of a 3x3 conv -> sub-pixel phases with pre-summed taps (c_network.py:215)
"""
import math

import torch

BN_EPS = 1e-5


def bn_affine(weight, bias, running_mean, running_covar, eps=BN_EPS):
    """Eval-mode ComplexBatchNorm2d as y = A [re; im] + c.  Returns (A (C,2,2), c (C,2)) in float64."""
    w, b = weight.detach().cpu().double(), bias.detach().cpu().double()
    cov = running_covar.detach().cpu().double()
    running_mean = running_mean.detach().cpu()
    mu = torch.stack([running_mean.real.double(), running_mean.imag.double()], dim=1)
    Crr, Cii, Cri = cov[:, 0] + eps, cov[:, 1] + eps, cov[:, 2]
    s = torch.sqrt(Crr * Cii - Cri * Cri)
    t = torch.sqrt(Crr + Cii + 2 * s)
    ist = 1.0 / (s * t)
    R = torch.stack([torch.stack([(Cii + s) * ist, -Cri * ist], 1), torch.stack([-Cri * ist, (Crr + s) * ist], 1)], 1)
    Wm = torch.stack([torch.stack([w[:, 0], w[:, 2]], 1), torch.stack([w[:, 2], w[:, 1]], 1)], 1)
    A = Wm @ R
    c = b - (A @ mu[:, :, None])[:, :, 0]
    return A, c


def bn_affine_from_sd(sd, prefix):
    return bn_affine(sd[prefix + "weight"], sd[prefix + "bias"], sd[prefix + "running_mean"], sd[prefix + "running_covar"])


def affine6(A, c):
    """(C,2,2),(C,2) -> (C,6) fp32 rows A00 A01 A10 A11 c0 c1 (dcs_cbn_apply / dcs_stft_fwd.bn_affine)."""
    return torch.cat([A.reshape(-1, 4), c], dim=1).float().contiguous()


def round_tf32(t):
    """fp32 -> nearest tf32 (10-bit mantissa, ties to even), returned as fp32: the tensor core truncates the low 13
    mantissa bits of fp32 operands, so weights are pre-rounded on the host."""
    i = t.detach().float().contiguous().view(torch.int32)
    i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def _phase_taps(k, up):
    """Rows of a k-tap 1-D kernel (pad k//2) applied after nearest x`up`: per phase, {source offset: [taps]}."""
    pad = k // 2
    out = []
    for ph in range(up):
        groups = {}
        for kk in range(k):
            d = math.floor((ph + kk - pad) / up)
            groups.setdefault(d, []).append(kk)
        out.append(sorted(groups.items()))
    return out


def _to_h16(t, dtype):
    """float64 -> fp16 / bf16 (round to nearest even); fp16 saturates at +-65504 like the kernels' stores."""
    if dtype == torch.float16:
        t = t.clamp(-65504.0, 65504.0)
    return t.to(dtype)


class PackedConv:
    """GEMM operands of one complex convolution layer (see include/dcsnet.h: dcs_cconv_params)."""

    def __init__(self, w_r, w_i, b_r=None, b_i=None, bn=None, transposed=False, stride=(1, 1), up=(1, 1),
                 act=0, device="cpu", tc_dtype=None, want_tf32=False, _block=None, want_bf16=False):
        # tc_dtype: torch.float16 / torch.bfloat16 = also pack the tcgen05 kind::f16 operand in that type (the type of the
        # activations it multiplies); want_bf16=True is the older spelling of tc_dtype=torch.bfloat16
        if want_bf16 and tc_dtype is None:
            tc_dtype = torch.bfloat16
        assert tc_dtype in (None, torch.float16, torch.bfloat16)
        self.tc_dtype = tc_dtype
        if _block is not None:      # PackedConv.from_real: the real block weight / bias are given directly
            M, bias = _block
            cout, _, cin, _, kh, kw = M.shape
        else:
            w_r, w_i = w_r.detach().double().cpu(), w_i.detach().double().cpu()
            if transposed:  # (Cin, Cout, k, k) -> equivalent conv weight (Cout, Cin, k, k), spatially flipped
                w_r = w_r.permute(1, 0, 2, 3).flip(2, 3)
                w_i = w_i.permute(1, 0, 2, 3).flip(2, 3)
            cout, cin, kh, kw = w_r.shape
            # real block weight M[(co,ro), (ci,ri), ky, kx]
            M = torch.zeros(cout, 2, cin, 2, kh, kw, dtype=torch.float64)
            M[:, 0, :, 0], M[:, 0, :, 1] = w_r, -w_i
            M[:, 1, :, 0], M[:, 1, :, 1] = w_i, w_r
            bias = torch.zeros(cout, 2, dtype=torch.float64)
            if b_r is not None:
                b_r, b_i = b_r.detach().double().cpu(), b_i.detach().double().cpu()
                bias[:, 0], bias[:, 1] = b_r - b_i, b_r + b_i
        if bn is not None:
            A, c = bn
            M = torch.einsum("oab,obcdyx->oacdyx", A, M)
            bias = torch.einsum("oab,ob->oa", A, bias) + c
        self.cin, self.cout, self.kh, self.kw = cin, cout, kh, kw
        self.stride, self.up, self.act = tuple(stride), tuple(up), act
        rows, cols = _phase_taps(kh, up[0]), _phase_taps(kw, up[1])
        self.phases = up[0] * up[1]
        self.ntaps = len(rows[0]) * len(cols[0])
        assert self.phases * self.ntaps <= 64
        dy, dx, mats = [], [], []
        for ph in range(up[0]):
            for pw in range(up[1]):
                for d_y, kys in rows[ph]:
                    for d_x, kxs in cols[pw]:
                        dy.append(d_y)
                        dx.append(d_x)
                        mats.append(M[:, :, :, :, kys][..., kxs].sum(dim=(4, 5)))  # (co,2,ci,2)
        self.dy, self.dx = dy, dx
        N, C2 = 2 * cout, 2 * cin
        self.n_pad = (N + 15) // 16 * 16
        Wt = torch.stack(mats, 0).reshape(self.phases, self.ntaps, N, C2)  # [p][t][n][k]
        Wp = torch.zeros(self.phases, self.ntaps, self.n_pad, C2, dtype=torch.float64)
        Wp[:, :, :N] = Wt
        self.w_ptnk = Wp                                          # [p][t][n_pad][C2] float64, CPU
        # FFMA operand: [p][t][k][n_pad] fp32 ;  tcgen05 operand: [p][n_pad][t*C2 + k] fp16 / bf16 (K contiguous)
        self.w_ffma = Wp.permute(0, 1, 3, 2).contiguous().float().to(device)
        self.w_tc = None
        if tc_dtype is not None:
            K = self.ntaps * C2
            self.k_pad = (K + 63) // 64 * 64  # one pipeline stage of the tcgen05 kernel = 64 16-bit values of K
            wt = torch.zeros(self.phases, self.n_pad, self.k_pad, dtype=torch.float64)
            wt[:, :, :K] = Wp.permute(0, 2, 1, 3).reshape(self.phases, self.n_pad, K)
            self.w_tc = _to_h16(wt, tc_dtype).contiguous().to(device)
        self.w_tc32 = None
        if want_tf32:
            K = self.ntaps * C2
            k_pad = (K + 31) // 32 * 32  # one pipeline stage = 128 bytes of K = 32 fp32
            wt = torch.zeros(self.phases, self.n_pad, k_pad, dtype=torch.float64)
            wt[:, :, :K] = Wp.permute(0, 2, 1, 3).reshape(self.phases, self.n_pad, K)
            self.w_tc32 = round_tf32(wt.float()).contiguous().to(device)
        b = torch.zeros(self.n_pad, dtype=torch.float64)
        b[:N] = bias.reshape(-1)
        self.bias = b.float().to(device)
        self.bias_host = b.float().tolist()
        self.w_tail = None
        if cout == 1 and tuple(up) == (2, 2) and cin == 16:
            # dcs_dec6_tail_fwd operand: [phase][tap][ci][(M00 M10 M01 M11)]  (M[n][ri], stored by columns)
            self.w_tail = Wp[:, :, :2].reshape(self.phases, self.ntaps, 2, cin, 2).permute(0, 1, 3, 4, 2) \
                .contiguous().float().to(device)


def real_bn_affine(weight, bias, running_mean, running_var, eps=BN_EPS):
    """Eval-mode torch.nn.BatchNorm2d over C real channels as the (A (C/2,2,2), c (C/2,2)) affine of C/2 channel PAIRS
    (diagonal A): the real path (r_network.py) stores real channels 2c, 2c+1 where the complex path stores (re, im) of
    channel c, so the same kernels and the same BN fold apply.  An odd C is padded with an identity channel."""
    s = weight.detach().double().cpu() / torch.sqrt(running_var.detach().double().cpu() + eps)
    t = bias.detach().double().cpu() - running_mean.detach().double().cpu() * s
    if s.numel() % 2:
        s, t = torch.cat([s, torch.ones(1, dtype=torch.float64)]), torch.cat([t, torch.zeros(1, dtype=torch.float64)])
    A = torch.zeros(s.numel() // 2, 2, 2, dtype=torch.float64)
    A[:, 0, 0], A[:, 1, 1] = s[0::2], s[1::2]
    return A, torch.stack([t[0::2], t[1::2]], dim=1)


def packed_conv_from_real(weight, bias=None, bn=None, transposed=False, **kw):
    """A REAL Conv2d / ConvTranspose2d (r_network.py:61-66, 89-103) as the operands of the complex-conv kernels: real
    channels are paired (2c, 2c+1) <-> (re, im) of pseudo-complex channel c, so the real weight W[co][ci] IS the block
    matrix M[(co/2, co%2), (ci/2, ci%2)] the kernels multiply by; odd channel counts (the 1-channel magnitude input, the
    1-channel mask output) are zero-padded.  `bn` = real_bn_affine(...) of the BatchNorm2d that follows."""
    W = weight.detach().double().cpu()
    if transposed:
        W = W.permute(1, 0, 2, 3).flip(2, 3)
    co, ci, kh, kw_ = W.shape
    cop, cip = co + co % 2, ci + ci % 2
    Wp = torch.zeros(cop, cip, kh, kw_, dtype=torch.float64)
    Wp[:co, :ci] = W
    M = Wp.reshape(cop // 2, 2, cip // 2, 2, kh, kw_)
    b = torch.zeros(cop, dtype=torch.float64)
    if bias is not None:
        b[:co] = bias.detach().double().cpu()
    return PackedConv(None, None, bn=bn, _block=(M, b.reshape(cop // 2, 2)), **kw)


STRIP_M = 128  # pixels (strip rows) per M tile of the row-strip kernel (csrc/cconv_strip.cu)


class StripConv:
    """Operands of dcs_cconv2d_strip_fwd (include/dcsnet.h) derived from a PackedConv: the layer flattened into a
    table of MMA items over row strips + the resident weight image.

    merged=True : all sub-pixel phases of a group form ONE accumulator (columns (ph, pw, n)), one MMA per
                  (source row dy, source, dx, 16-channel K slice) with zero blocks where a phase has no such tap
                  (fewer, wider MMAs: the A operand read from shared memory is what bounds small-N MMAs).
    merged=False: one MMA per (phase, pre-summed tap, source, K slice) with N = 2*cout.
    groups      : 1, or 2 = one phase row (ph) per CTA group (halves the resident weights and the ring depth).
    """

    def __init__(self, pk, c0, c1, merged=True, groups=1, device="cpu", dtype=None):
        assert c0 + c1 == pk.cin
        dtype = dtype or pk.tc_dtype or torch.bfloat16
        uh, uw = pk.up
        sh, sw = pk.stride
        N = 2 * pk.cout
        assert N & (N - 1) == 0 and N >= 8
        assert groups in (1, 2) and (groups == 1 or uh == 2)
        self.pk, self.c0, self.c1 = pk, c0, c1
        self.up, self.stride = tuple(pk.up), tuple(pk.stride)
        Wp = pk.w_ptnk
        T = pk.ntaps
        dy = [[pk.dy[p * T + t] for t in range(T)] for p in range(pk.phases)]
        dx = [[pk.dx[p * T + t] for t in range(T)] for p in range(pk.phases)]
        Pb = [4 * c0, 4 * c1]                              # bytes of one pixel per source (bf16 complex)
        P = [b * sw for b in Pb]                           # strip row bytes (pixel pairs when stride_w = 2)
        for s in range(2 if c1 else 1):
            assert P[s] in (32, 64, 128), f"strip rows must be 32/64/128 bytes, got {P[s]}"
        unit = lambda d: d // sw                           # strip row (floor) and parity of a tap offset
        half = lambda d: d % sw
        all_dx = sorted({d for row in dx for d in row})
        self.x_min = min(unit(d) for d in all_dx)
        self.box_units = STRIP_M + max(unit(d) for d in all_dx) - self.x_min
        strip_off = [0, (self.box_units * P[0] + 1023) // 1024 * 1024]
        koff = [0, 2 * c0]
        run = uw * N
        ph_sets = [list(range(uh))] if groups == 1 else [[0], [1]]
        self.cols = len(ph_sets[0]) * run
        assert self.cols in (16, 32, 64), f"accumulator columns per row tile must be 16/32/64, got {self.cols}"
        self.n_mma = (self.cols + 15) // 16 * 16 if merged else N
        assert self.n_mma % 16 == 0 and self.n_mma <= self.cols
        items, blocks, self.groups = [], [], []
        w_off = 0
        for phs in ph_sets:
            plist = [(pl, ph, pw) for pl, ph in enumerate(phs) for pw in range(uw)]
            dys = sorted({dy[ph * uw + pw][t] for _, ph, pw in plist for t in range(T)})
            dy_min, n_dy = dys[0], dys[-1] - dys[0] + 1
            g_items, g_blocks = [], []

            def add(drow, s, d_x, ks, d_col, blk):
                a_off = strip_off[s] + (unit(d_x) - self.x_min) * P[s] + half(d_x) * Pb[s] + ks * 32
                g_items.append(dict(a_off16=a_off // 16, b_off16=len(g_blocks) * self.n_mma * 32 // 16, d_col=d_col,
                                    drow=drow, src=s))
                g_blocks.append(blk)

            for s in range(2 if c1 else 1):
                assert Pb[s] % 32 == 0
            if merged:
                for d_y in dys:
                    for s in range(2 if c1 else 1):
                        for d_x in all_dx:
                            taps = [(pl, pw, ph * uw + pw, t) for pl, ph, pw in plist for t in range(T)
                                    if dy[ph * uw + pw][t] == d_y and dx[ph * uw + pw][t] == d_x]
                            if not taps:
                                continue
                            for ks in range(Pb[s] // 32):
                                blk = torch.zeros(self.n_mma, 16, dtype=torch.float64)
                                k0 = koff[s] + ks * 16
                                for pl, pw, p, t in taps:
                                    c = (pl * uw + pw) * N
                                    blk[c:c + N] += Wp[p, t, :N, k0:k0 + 16]
                                add(d_y - dy_min, s, d_x, ks, 0, blk)
            else:
                for pl, ph, pw in plist:
                    p = ph * uw + pw
                    for t in range(T):
                        for s in range(2 if c1 else 1):
                            for ks in range(Pb[s] // 32):
                                k0 = koff[s] + ks * 16
                                add(dy[p][t] - dy_min, s, dx[p][t], ks, (pl * uw + pw) * N, Wp[p, t, :N, k0:k0 + 16].clone())
                order = sorted(range(len(g_items)), key=lambda i: g_items[i]["drow"])   # stable: oldest ring rows first
                g_items = [g_items[i] for i in order]
            seen = set()
            for it in g_items:
                it["first"] = it["d_col"] not in seen
                seen.add(it["d_col"])
            w_bytes = len(g_blocks) * self.n_mma * 32
            self.groups.append(dict(item0=len(items), n_items=len(g_items), dy_min=dy_min, n_dy=n_dy, ph0=phs[0],
                                    n_ph=len(phs), x_min=self.x_min, w_bytes=w_bytes, w_off=w_off))
            w_off += (w_bytes + 1023) // 1024 * 1024
            items += g_items
            blocks.append(g_blocks)
        self.items = items
        self.item_table = _strip_item_table(items)
        self.w_image = _strip_weight_image(self.groups, blocks, self.n_mma, w_off, dtype).to(device)
        self.smem_weight_bytes = max(g["w_bytes"] for g in self.groups)


def _strip_item_table(items):
    """uint32 x 4 per item {a_off16, b_off16, d_col | drow << 16 | flags << 24, 0}; HOST table (kernel parameters)."""
    tab = torch.zeros(len(items), 4, dtype=torch.int64)
    for i, it in enumerate(items):
        flags = (1 if it["first"] else 0) | (2 if it["src"] else 0)
        tab[i, 0], tab[i, 1] = it["a_off16"], it["b_off16"]
        tab[i, 2] = it["d_col"] | (it["drow"] << 16) | (flags << 24)
    return tab.to(torch.int32).contiguous()


def _strip_weight_image(groups, blocks, n_mma, total_bytes, dtype=torch.bfloat16):
    """Per group, blocks of [n_mma][16] fp16 / bf16 as rows of 32 bytes with the SWIZZLE_32B pattern (the 16-byte halves of a row
    swap when bit 7 of the byte offset is set): the exact shared-memory image, bulk-copied by the kernel."""
    img = torch.zeros(total_bytes // 2, dtype=dtype)
    n = torch.arange(n_mma)[:, None]
    k = torch.arange(16)[None, :]
    off = n * 32 + k * 2
    off = off ^ (((off >> 7) & 1) << 4)
    for g, g_blocks in zip(groups, blocks):
        for bi, blk in enumerate(g_blocks):
            img[(g["w_off"] + bi * n_mma * 32 + off) // 2] = _to_h16(blk, dtype)
    return img.contiguous()


class StripEnc0:
    """encoder[0] (ComplexConv2d 1 -> 8, k7, stride (2,2), on the initial_batchnorm output) for dcs_cconv2d_strip_fwd.

    One complex input channel gives K = 2 per tap, far below an MMA's K = 16, so K is taken from SPACE: a strip row is
    16 consecutive source pixels (64 bytes) and the 8 output pixels whose windows start in it form N = 8 x 16 columns;
    a K slice is 8 source pixels, and each (kernel row dy, slice) item multiplies by a Toeplitz block holding the taps
    dx = 8 (s - 1) + p - 2 j  (source pixel p of slice s, output pixel j).  4 slices x 7 rows = 28 MMAs (N = 128) per
    1024 output pixels.  The kernel is told a reinterpreted geometry: source (B, F, T/8, 8 "channels"), stride_w 2
    (a strip row = 2 such pixels), up_w = 8 (8 output pixels per strip row), cout 8."""

    def __init__(self, pk, device="cpu", dtype=None):
        assert (pk.cin, pk.cout, pk.kh, pk.kw, tuple(pk.stride), tuple(pk.up)) == (1, 8, 7, 7, (2, 2), (1, 1))
        dtype = dtype or pk.tc_dtype or torch.bfloat16
        self.pk, self.c0, self.c1 = pk, 8, 0
        self.up, self.stride = (1, 8), (2, 2)
        N = 16
        Wp = pk.w_ptnk                                     # [1][49][16][2]
        tap = {(pk.dy[t], pk.dx[t]): t for t in range(pk.ntaps)}
        self.x_min, self.box_units, self.cols, self.n_mma = -1, STRIP_M + 2, 8 * N, 8 * N
        shift, koff = [0, 1, 1, 2], [32, 0, 32, 0]        # slice s: strip row v - 1 + shift[s], byte offset in the row
        items, blocks = [], []
        for d_y in range(-3, 4):
            for sl in range(4):
                blk = torch.zeros(self.n_mma, 16, dtype=torch.float64)
                for j in range(8):
                    for p_ in range(8):
                        d_x = 8 * (sl - 1) + p_ - 2 * j
                        if -3 <= d_x <= 3:
                            blk[j * N:(j + 1) * N, 2 * p_:2 * p_ + 2] = Wp[0, tap[(d_y, d_x)], :N, :]
                items.append(dict(a_off16=(shift[sl] * 64 + koff[sl]) // 16, b_off16=len(blocks) * self.n_mma * 32 // 16,
                                  d_col=0, drow=d_y + 3, src=0, first=not items))
                blocks.append(blk)
        w_bytes = len(blocks) * self.n_mma * 32
        self.groups = [dict(item0=0, n_items=len(items), dy_min=-3, n_dy=7, ph0=0, n_ph=1, x_min=self.x_min,
                            w_bytes=w_bytes, w_off=0)]
        self.items = items
        self.item_table = _strip_item_table(items)
        self.w_image = _strip_weight_image(self.groups, [blocks], self.n_mma, w_bytes, dtype).to(device)
        self.smem_weight_bytes = w_bytes

    @staticmethod
    def view_src(bn0):
        """(B, F, T, 1, 2) bf16 initial_batchnorm output -> the (B, F, T/8, 8, 2) view the kernel is given."""
        B, F, T = bn0.shape[:3]
        assert T % 16 == 0
        return bn0.view(B, F, T // 8, 8, 2)


class StripDec6:
    """decoder[6] (ComplexConvTranspose2d 16 -> 1 after cat + (2,2) up-sampling) for dcs_cconv2d_strip_fwd with the fused
    mask tail.  N = 2 real outputs per pixel is far below an MMA's minimum, so N is taken from SPACE: a strip row is 4
    source pixels (128 bytes per source), a K slice is one source pixel (8 complex channels), and the accumulator
    columns are (phase row ph, source pixel j, phase column pw, re/im) = 2 x 16: every (dy, source, pixel offset q) item
    multiplies by a Toeplitz block with the pre-summed taps dx = q - j.  6 offsets x 2 sources x 3 rows = 36 MMAs
    (N = 32) per 512 source pixels.  Kernel geometry: sources viewed as (B, H, W/4, 32 "channels"), up = (2, 8)."""

    def __init__(self, pk, device="cpu", dtype=None):
        assert (pk.cin, pk.cout, tuple(pk.up), tuple(pk.stride)) == (16, 1, (2, 2), (1, 1))
        dtype = dtype or pk.tc_dtype or torch.bfloat16
        self.pk, self.c0, self.c1 = pk, 32, 32
        self.up, self.stride = (2, 8), (1, 1)
        Wp, T = pk.w_ptnk, pk.ntaps                       # [4 phases][4 taps][n_pad][32]
        tap = [{(pk.dy[p * T + t], pk.dx[p * T + t]): t for t in range(T)} for p in range(4)]
        self.x_min, self.box_units, self.cols, self.n_mma = -1, STRIP_M + 2, 32, 32
        strip_off = [0, (self.box_units * 128 + 1023) // 1024 * 1024]
        items, blocks = [], []
        for d_y in (-1, 0, 1):
            for src in range(2):
                for q in range(-1, 5):                     # source pixel offset relative to the strip row's first pixel
                    blk = torch.zeros(32, 16, dtype=torch.float64)
                    for j in range(4):
                        d_x = q - j
                        for ph in range(2):
                            for pw in range(2):
                                t = tap[ph * 2 + pw].get((d_y, d_x))
                                if t is not None:
                                    c = ph * 16 + j * 4 + pw * 2
                                    blk[c:c + 2] = Wp[ph * 2 + pw, t, :2, src * 16:(src + 1) * 16]
                    shift, ks = (0, 3) if q < 0 else ((2, 0) if q > 3 else (1, q))
                    items.append(dict(a_off16=(strip_off[src] + shift * 128 + ks * 32) // 16, b_off16=len(blocks) * 32 * 32 // 16,
                                      d_col=0, drow=d_y + 1, src=src, first=not items))
                    blocks.append(blk)
        w_bytes = len(blocks) * 32 * 32
        self.groups = [dict(item0=0, n_items=len(items), dy_min=-1, n_dy=3, ph0=0, n_ph=2, x_min=self.x_min,
                            w_bytes=w_bytes, w_off=0)]
        self.items = items
        self.item_table = _strip_item_table(items)
        self.w_image = _strip_weight_image(self.groups, [blocks], 32, w_bytes, dtype).to(device)
        self.smem_weight_bytes = w_bytes

    @staticmethod
    def view_src(x):
        """(B, H, W, 8, 2) bf16 -> the (B, H, W/4, 32, 2) view the kernel is given (W % 4 == 0)."""
        B, H, W = x.shape[:3]
        assert W % 4 == 0
        return x.view(B, H, W // 4, 32, 2)


def pack_lstm(sd, prefix, device, hidden=64, layers=2):
    """ComplexLSTM weights (c_network.py:17-20) -> (w_ih0 [D][1024], w_ih1 [2][128][512], w_hh [2][2][2][256][64],
    bias [1024 + 2*512]) as consumed by dcs_clstm_fwd.  Column order n = lstm*512 + dir*256 + gate."""
    assert layers == 2
    names = ("real_lstm", "imag_lstm")
    sfx = ("", "_reverse")
    g = lambda k: sd[prefix + k].detach().double().cpu()
    w0 = torch.cat([g(f"{n}.weight_ih_l0{s}") for n in names for s in sfx], 0)           # (1024, D)
    b0 = torch.cat([g(f"{n}.bias_ih_l0{s}") + g(f"{n}.bias_hh_l0{s}") for n in names for s in sfx], 0)
    w1 = torch.stack([torch.cat([g(f"{n}.weight_ih_l1{s}") for s in sfx], 0).t() for n in names], 0)  # (2,128,512)
    b1 = torch.cat([g(f"{n}.bias_ih_l1{s}") + g(f"{n}.bias_hh_l1{s}") for n in names for s in sfx], 0)
    whh = torch.stack([torch.stack([torch.stack([g(f"{n}.weight_hh_l{l}{s}") for s in sfx], 0) for n in names], 0)
                       for l in range(2)], 0)                                                # (2,2,2,256,64)
    f = lambda t: t.contiguous().float().to(device)
    # tensor-core operands: K-major slices [lstm*2+dir][4H][K], pre-rounded to tf32
    w0_t = round_tf32(w0.float().reshape(4, 4 * hidden, -1))
    w1_t = round_tf32(w1.permute(0, 2, 1).float().reshape(4, 4 * hidden, 2 * hidden))
    return dict(w_ih0=f(w0.t()), w_ih1=f(w1), w_hh=f(whh), bias=f(torch.cat([b0, b1], 0)),
                w_ih0_t=f(w0_t), w_ih1_t=f(w1_t), w_hh_frag=f(lstm_whh_fragments(whh.float(), hidden)))


def lstm_whh_fragments(whh, hidden=64):
    """W_hh (layer, lstm, dir, 4H, H) -> tf32 A fragments of mma.sync.m16n8k8 for lstm_recurrent4_mma_kernel:
    [layer][lstm][dir][warp 8][tile 2][kstep 8][lane 32][4].  Warp w owns units 8w..8w+7; tile 0 rows 0-7 / 8-15 are the
    gates (i, f) of those units, tile 1 the gates (g, o); fragment register j of lane l is row l/4 + 8 (j & 1),
    column l%4 + 4 (j >> 1) of the 16 x 8 tile of k-step ks (logical k = 8 ks + column)."""
    H = hidden
    warp = torch.arange(8).view(8, 1, 1, 1, 1)
    tile = torch.arange(2).view(1, 2, 1, 1, 1)
    ks = torch.arange(8).view(1, 1, 8, 1, 1)
    lane = torch.arange(32).view(1, 1, 1, 32, 1)
    j = torch.arange(4).view(1, 1, 1, 1, 4)
    row = lane // 4 + 8 * (j & 1)
    gate = 2 * tile + (row >= 8).long()
    unit = 8 * warp + row % 8
    k = 8 * ks + lane % 4 + 4 * (j >> 1)
    r = (gate * H + unit).expand(8, 2, 8, 32, 4)
    c = k.expand(8, 2, 8, 32, 4)
    return round_tf32(whh[:, :, :, r, c]).contiguous()


def pack_channel_attention(sd, prefix, device):
    f = lambda k: sd[prefix + k].detach().float().reshape(sd[prefix + k].shape[0], -1).contiguous().to(device)
    d = dict(w1_r=f("fc.0.conv_r.weight"), w1_i=f("fc.0.conv_i.weight"),
             w2_r=f("fc.2.conv_r.weight"), w2_i=f("fc.2.conv_i.weight"))
    d["reduced"], d["channels"] = d["w1_r"].shape
    return d


def pack_spatial_attention(sd, prefix, device):
    wr, wi = sd[prefix + "conv1.conv_r.weight"], sd[prefix + "conv1.conv_i.weight"]
    assert tuple(wr.shape) == (1, 2, 7, 7), "only spatial_attention_kernel_size = 7 is built"
    return torch.cat([wr.detach().float().reshape(-1), wi.detach().float().reshape(-1)]).contiguous().to(device)


class PackedRNet:
    """GEMM-ready operands of an R_NETWORK state_dict (r_network.py) in the layout of the complex path:
    real channels (2c, 2c+1) are the (re, im) of pseudo-complex channel c, so the convs, the fc and the initial BatchNorm
    run on the existing kernels; the real CBAM / LSTM weights are kept in the reference's own layout for the round-2
    kernels.  Host-only so far (tests/emulate.rnet_dataflow evaluates the network through these operands on the CPU)."""

    KERNEL_E = [7, 7, 5, 5, 3, 3, 3]
    STRIDE_E = [(2, 2), (2, 2), (2, 2), (2, 1), (2, 1), (2, 1), (2, 1)]
    UPSAMPLE = [(2, 1), (2, 1), (2, 1), (2, 1), (2, 2), (2, 2), (2, 2)]

    def __init__(self, model_or_sd, device="cpu", no_of_layers=7, want_bf16=False, tc_dtype=None):
        if want_bf16 and tc_dtype is None:
            tc_dtype = torch.bfloat16
        self.tc_dtype = tc_dtype
        sd = model_or_sd.state_dict() if hasattr(model_or_sd, "state_dict") else model_or_sd
        sd = {k: v.detach() for k, v in sd.items()}
        Lr = self.L = no_of_layers
        bn = lambda p: real_bn_affine(sd[p + "weight"], sd[p + "bias"], sd[p + "running_mean"], sd[p + "running_var"])  # noqa: E731
        self.bn0 = affine6(*bn("initial_batchnorm.")).to(device)          # magnitude in .re, 0 in .im (identity channel)
        self.enc, self.dec, self.skip_att, self.dec_att = [], [], [], []
        att = lambda p: dict(w1=sd[p + "fc.0.weight"].float().flatten(1).contiguous().to(device),   # noqa: E731
                             w2=sd[p + "fc.2.weight"].float().flatten(1).contiguous().to(device))
        for i in range(Lr):
            p = f"encoder.{i}."
            self.enc.append(packed_conv_from_real(sd[p + "0.weight"], sd[p + "0.bias"], bn=bn(p + "1."), stride=self.STRIDE_E[i],
                                                  act=1, device=device, tc_dtype=tc_dtype))
        for i in range(Lr):
            last = i == Lr - 1
            p = f"decoder.{i}." if last else f"decoder.{i}.0."
            self.dec.append(packed_conv_from_real(sd[p + "weight"], sd[p + "bias"], bn=None if last else bn(f"decoder.{i}.1."),
                                                  transposed=True, up=self.UPSAMPLE[i], act=3 if last else 2, device=device,   # last: torch.sigmoid (r_network.py:172)
                                                  tc_dtype=tc_dtype))
            self.skip_att.append((att(f"skip_attention.{2 * i}."), sd[f"skip_attention.{2 * i + 1}.conv1.weight"].float().reshape(2, 49).contiguous().to(device)))
            if not last:
                self.dec_att.append((att(f"decoder_attention.{2 * i}."), sd[f"decoder_attention.{2 * i + 1}.conv1.weight"].float().reshape(2, 49).contiguous().to(device)))
        self.fc = packed_conv_from_real(sd["fc.weight"][:, :, None, None], sd["fc.bias"], device=device, tc_dtype=tc_dtype)
        # nn.LSTM(256 -> 128, 2 layers, bidirectional): [layer][dir] -> (w_ih (4H, in), w_hh (4H, H), b_ih + b_hh)
        self.lstm = []
        for layer in range(2):
            dirs = []
            for suf in ("", "_reverse"):
                k = f"lstm.{{}}_l{layer}{suf}"
                dirs.append(dict(w_ih=sd[k.format("weight_ih")].float().contiguous().to(device),
                                 w_hh=sd[k.format("weight_hh")].float().contiguous().to(device),
                                 bias=(sd[k.format("bias_ih")] + sd[k.format("bias_hh")]).float().contiguous().to(device)))
            self.lstm.append(dirs)
        # tensor-core operands (dcs_rlstm_tc_fwd): input projections [dir][half][256 gate rows][in] in the 16-bit storage type,
        # W_hh fp32 [layer][dir][4H][H] (the kernel builds its fp16 register fragments itself), bias fp32 [layer][dir][4H]
        self.lstm_tc = None
        if tc_dtype is not None:
            ih = lambda layer: torch.stack([d["w_ih"].reshape(2, 256, -1) for d in self.lstm[layer]]).to(tc_dtype).contiguous()   # noqa: E731
            self.lstm_tc = dict(w_ih0=ih(0), w_ih1=ih(1),
                                w_hh=torch.stack([torch.stack([d["w_hh"] for d in self.lstm[layer]]) for layer in range(2)]).contiguous(),
                                bias=torch.stack([torch.stack([d["bias"] for d in self.lstm[layer]]) for layer in range(2)]).contiguous())
        # the same weights transposed for dcs_rlstm_fwd: w_ih*_t [in][2*4H] (columns dir*4H + gate row), w_hh_t [layer][dir][H][4H]
        cat_t = lambda layer, k: torch.cat([d[k].t() for d in self.lstm[layer]], dim=1).contiguous()   # noqa: E731
        self.lstm_t = dict(w_ih0_t=cat_t(0, "w_ih"), w_ih1_t=cat_t(1, "w_ih"),
                           w_hh_t=torch.stack([torch.stack([d["w_hh"].t().contiguous() for d in self.lstm[layer]]) for layer in range(2)]).contiguous(),
                           bias=torch.stack([torch.cat([d["bias"] for d in self.lstm[layer]]) for layer in range(2)]).contiguous())
