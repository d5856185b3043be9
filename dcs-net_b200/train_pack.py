"""Operand re-packing for the training step, ON THE GPU, one launch per optimizer step.

packing.PackedConv builds the conv kernels' operands on the host in float64 — fine once per checkpoint for inference, not once
per optimizer step.  Every operand element is a signed sum of at most four raw parameter elements (block matrix [[Wr, -Wi], [Wi,
Wr]], transposed-conv flip, sub-pixel phase pre-sums of up to 2 x 2 taps, bias rule (b_r - b_i, b_r + b_i), role swaps of the
dgrad operands, LSTM weight stacking / bias sums), so the whole map raw parameters -> operands is a sparse {-1, 0, +1} matrix with
<= 4 entries per row.  `Sym` re-runs the packing algebra on (index, sign) tensors instead of values; the resulting tables drive
dcs_gather_pack (csrc/train_bwd.cu), which rebuilds every operand from the flat fp32 parameter buffer.

Reference: the layouts are those of packing.PackedConv / train_ops.dgrad_conv (c_network.py:107-112, 135-147; complexPyTorch
apply_complex); tests/test_train_pack.py checks the tables against those packers on the CPU.
"""
import torch

from . import _lib as L
from .packing import _phase_taps

MAX_TERMS = 4


class Sym:
    """A tensor whose elements are signed sums of flat-parameter elements: idx (..., T) int64 (-1 = no term), sgn (..., T) int8."""

    def __init__(self, idx, sgn):
        self.idx, self.sgn = idx, sgn

    @classmethod
    def leaf(cls, offset, shape):
        n = 1
        for s in shape:
            n *= s
        return cls((offset + torch.arange(n, dtype=torch.int64)).view(*shape, 1), torch.ones(*shape, 1, dtype=torch.int8))

    @classmethod
    def zeros(cls, *shape):
        return cls(torch.full((*shape, 1), -1, dtype=torch.int64), torch.zeros(*shape, 1, dtype=torch.int8))

    @property
    def shape(self):
        return tuple(self.idx.shape[:-1])

    def __neg__(self):
        return Sym(self.idx, -self.sgn)

    def _map(self, f):
        return Sym(f(self.idx), f(self.sgn))

    def permute(self, *d):
        return self._map(lambda t: t.permute(*d, len(d)))

    def flip(self, *dims):
        return self._map(lambda t: t.flip(*dims))

    def reshape(self, *shape):
        return self._map(lambda t: t.reshape(*shape, t.shape[-1]))

    def index_select(self, dim, ids):
        i = torch.as_tensor(ids, dtype=torch.int64)
        return self._map(lambda t: t.index_select(dim, i))

    def __getitem__(self, key):
        return self._map(lambda t: t[key])

    def sum(self, dims):
        """Sum over leading dims `dims`: their elements become additional terms."""
        nd = len(self.shape)
        keep = [d for d in range(nd) if d not in dims]
        def f(t):
            t = t.permute(*keep, *dims, nd)
            return t.reshape(*t.shape[:len(keep)], -1)
        return self._map(f)

    def pad_terms(self, T):
        t = self.idx.shape[-1]
        if t == T:
            return self
        pad = (*self.shape, T - t)
        return Sym(torch.cat([self.idx, torch.full(pad, -1, dtype=torch.int64)], -1), torch.cat([self.sgn, torch.zeros(pad, dtype=torch.int8)], -1))

    def __add__(self, other):
        return Sym(torch.cat([self.idx, other.idx], -1), torch.cat([self.sgn, other.sgn], -1))

    def __sub__(self, other):
        return self + (-other)

    @staticmethod
    def _same_terms(items):
        T = max(s.idx.shape[-1] for s in items)
        return [s.pad_terms(T) for s in items]

    @staticmethod
    def stack(items, dim):
        items = Sym._same_terms(items)
        return Sym(torch.stack([s.idx for s in items], dim), torch.stack([s.sgn for s in items], dim))

    @staticmethod
    def cat(items, dim):
        items = Sym._same_terms(items)
        return Sym(torch.cat([s.idx for s in items], dim), torch.cat([s.sgn for s in items], dim))

    def tables(self):
        """(n, 4) int32 indices and (n, 4) int8 signs, row-major over the tensor's elements."""
        idx, sgn = self.idx.reshape(-1, self.idx.shape[-1]), self.sgn.reshape(-1, self.sgn.shape[-1])
        order = torch.argsort((idx < 0).to(torch.int8), dim=1, stable=True)          # real terms first
        idx, sgn = torch.gather(idx, 1, order), torch.gather(sgn, 1, order)
        if idx.shape[1] > MAX_TERMS:
            assert bool((idx[:, MAX_TERMS:] < 0).all()), "an operand element has more than four terms"
            idx, sgn = idx[:, :MAX_TERMS], sgn[:, :MAX_TERMS]
        elif idx.shape[1] < MAX_TERMS:
            pad = MAX_TERMS - idx.shape[1]
            idx = torch.cat([idx, torch.full((idx.shape[0], pad), -1, dtype=torch.int64)], 1)
            sgn = torch.cat([sgn, torch.zeros(sgn.shape[0], pad, dtype=torch.int8)], 1)
        sgn = torch.where(idx < 0, torch.zeros_like(sgn), sgn)
        return idx.to(torch.int32).contiguous(), sgn.contiguous()

    def evaluate(self, flat):
        """Numeric value on a flat parameter vector (CPU check of the tables)."""
        idx, sgn = self.tables()
        vals = flat.double()[idx.clamp_min(0).long()] * sgn.double()
        return vals.sum(1).reshape(self.shape)


def sym_conv(w_r, w_i, b_r=None, b_i=None, transposed=False, up=(1, 1), tc=False, tf32=False):
    """packing.PackedConv's operands from symbolic raw weights: dict(w_ffma [p][t][k][n_pad], bias [n_pad], w_tc [p][n_pad][k_pad])."""
    if transposed:
        w_r, w_i = w_r.permute(1, 0, 2, 3).flip(2, 3), w_i.permute(1, 0, 2, 3).flip(2, 3)
    cout, cin, kh, kw = w_r.shape
    M = Sym.stack([Sym.stack([w_r, -w_i], 2), Sym.stack([w_i, w_r], 2)], 1)               # (co, ro, ci, ri, ky, kx)
    rows, cols = _phase_taps(kh, up[0]), _phase_taps(kw, up[1])
    mats = []
    for ph in range(up[0]):
        for pw in range(up[1]):
            for _, kys in rows[ph]:
                for _, kxs in cols[pw]:
                    mats.append(M.index_select(4, kys).index_select(5, kxs).sum((4, 5)))    # (co, 2, ci, 2)
    phases, N, C2 = up[0] * up[1], 2 * cout, 2 * cin
    ntaps = len(mats) // phases
    n_pad = (N + 15) // 16 * 16
    Wt = Sym.stack(mats, 0).reshape(phases, ntaps, N, C2)
    if n_pad > N:
        Wt = Sym.cat([Wt, Sym.zeros(phases, ntaps, n_pad - N, C2)], 2)
    out = dict(w_ffma=Wt.permute(0, 1, 3, 2))
    if tc:
        K = ntaps * C2
        k_pad = (K + 63) // 64 * 64
        wt = Wt.permute(0, 2, 1, 3).reshape(phases, n_pad, K)
        if k_pad > K:
            wt = Sym.cat([wt, Sym.zeros(phases, n_pad, k_pad - K)], 2)
        out["w_tc"] = wt
    if tf32:
        K = ntaps * C2
        k_pad = (K + 31) // 32 * 32
        wt = Wt.permute(0, 2, 1, 3).reshape(phases, n_pad, K)
        if k_pad > K:
            wt = Sym.cat([wt, Sym.zeros(phases, n_pad, k_pad - K)], 2)
        out["w_tc32"] = wt
    if b_r is not None:
        b = Sym.stack([b_r - b_i, b_r + b_i], 1).reshape(N)
        if n_pad > N:
            b = Sym.cat([b, Sym.zeros(n_pad - N)], 0)
        out["bias"] = b
    return out


def sym_dgrad(w_r, w_i, transposed, tc=False, tf32=False):
    """train_ops.dgrad_conv's operands (role-swapped weights of a stride-1 layer's data gradient)."""
    if transposed:
        return sym_conv(w_r, -w_i, tc=tc, tf32=tf32)
    return sym_conv(w_r.permute(1, 0, 2, 3).flip(2, 3), -(w_i.permute(1, 0, 2, 3).flip(2, 3)), tc=tc, tf32=tf32)


def sym_dgrad_strided(w_r, w_i, stride, tf32=False):
    """train_ops.PhasePack's operands (the data gradient of a strided conv as a phase convolution) from symbolic raw weights."""
    from .train_ops import strided_dgrad_taps
    cout, cin, kh, kw = w_r.shape
    wr, wi = w_r.permute(1, 0, 2, 3), w_i.permute(1, 0, 2, 3)
    M = Sym.stack([Sym.stack([wr, wi], 2), Sym.stack([-wi, wr], 2)], 1)                    # (ci, ro, co, ri, ky, kx)
    rows, cols = strided_dgrad_taps(kh, stride[0]), strided_dgrad_taps(kw, stride[1])
    mats = []
    for ph in range(stride[0]):
        for pw in range(stride[1]):
            for _, ky in rows[ph]:
                for _, kx in cols[pw]:
                    mats.append(Sym.zeros(cin, 2, cout, 2) if ky is None or kx is None else M[:, :, :, :, ky, kx])
    phases, N, C2 = stride[0] * stride[1], 2 * cin, 2 * cout
    ntaps = len(mats) // phases
    n_pad = (N + 15) // 16 * 16
    Wt = Sym.stack(mats, 0).reshape(phases, ntaps, N, C2)
    if n_pad > N:
        Wt = Sym.cat([Wt, Sym.zeros(phases, ntaps, n_pad - N, C2)], 2)
    out = dict(w_ffma=Wt.permute(0, 1, 3, 2))
    if tf32:
        K = ntaps * C2
        k_pad = (K + 31) // 32 * 32
        wt = Wt.permute(0, 2, 1, 3).reshape(phases, n_pad, K)
        if k_pad > K:
            wt = Sym.cat([wt, Sym.zeros(phases, n_pad, k_pad - K)], 2)
        out["w_tc32"] = wt
    return out


class GatherPack:
    """Collects (symbolic operand, destination tensor) pairs and rebuilds all destinations from the flat parameter buffer with one
    dcs_gather_pack launch per destination dtype."""

    def __init__(self, flat_param):
        self.flat = flat_param
        self.jobs = {}          # dtype -> list of (Sym, shape, setter)
        self.current = []

    def add(self, sym, dtype, setter, current=None):
        """dtype: torch.float32 / float16 / bfloat16, or "tf32" (fp32 storage rounded to tf32).  `current`: the host-packed tensor this
        job replaces (kept for `check_against_host_packing`)."""
        self.jobs.setdefault(dtype, []).append((sym, setter))
        self.current.append((sym, dtype, current))

    def check_against_host_packing(self, flat=None):
        """Every table evaluated on the flat parameter vector (on the CPU) equals the tensor the host packing produced — the whole-model
        form of tests/test_train_pack.py.  Returns the number of operands compared."""
        from .packing import round_tf32
        flat = (self.flat if flat is None else flat).detach().cpu()
        n = 0
        for sym, dtype, cur in self.current:
            if cur is None:
                continue
            got = sym.evaluate(flat).float()
            if dtype == "tf32":
                got = round_tf32(got)
            elif dtype in (torch.float16, torch.bfloat16):
                got = got.to(dtype).float()
            want = cur.detach().cpu().float()
            assert tuple(got.shape) == tuple(want.shape), (tuple(got.shape), tuple(want.shape))
            assert torch.allclose(got, want, rtol=1e-2 if dtype in (torch.float16, torch.bfloat16) else 1e-6, atol=1e-7), dtype
            n += 1
        return n

    def add_packed_conv(self, pk, sy):
        """Redirect the operand tensors of a packing.PackedConv to gather-pack outputs (sy = sym_conv / sym_dgrad of the same layer)."""
        for name, dtype in (("w_ffma", torch.float32), ("bias", torch.float32), ("w_tc32", "tf32"), ("w_tc", getattr(pk, "tc_dtype", None))):
            if dtype is None:
                continue
            cur = getattr(pk, name, None)
            if cur is None or name not in sy:
                continue
            assert tuple(cur.shape) == tuple(sy[name].shape), (name, tuple(cur.shape), tuple(sy[name].shape))
            self.add(sy[name], dtype, lambda t, pk=pk, name=name: setattr(pk, name, t), current=cur)

    def finalize(self):
        dev = self.flat.device
        self.plans = []
        for dtype, jobs in self.jobs.items():
            tabs = [s.tables() for s, _ in jobs]
            sizes = [t[0].shape[0] for t in tabs]
            # 16-byte aligned slots so that every operand keeps the alignment the kernels' vector loads assume
            offs, total = [], 0
            for n in sizes:
                offs.append(total)
                total += (n + 63) // 64 * 64
            idx = torch.full((total, 4), -1, dtype=torch.int32)
            sgn = torch.zeros((total, 4), dtype=torch.int8)
            for (ti, ts), o, n in zip(tabs, offs, sizes):
                idx[o:o + n], sgn[o:o + n] = ti, ts
            out = torch.zeros(total, dtype=torch.float32 if dtype == "tf32" else dtype, device=dev)
            for (s, setter), o, n in zip(jobs, offs, sizes):
                setter(out[o:o + n].view(s.shape))
            self.plans.append((idx.to(dev), sgn.to(dev), out, total, 3 if dtype == "tf32" else L.dtype_code(out)))
        return self

    def run(self):
        for idx, sgn, out, n, code in self.plans:
            L.check(L.lib().dcs_gather_pack(L.ptr(self.flat), L.ptr(idx), L.ptr(sgn), L.ptr(out), n, code, L.stream_ptr()), "dcs_gather_pack")
