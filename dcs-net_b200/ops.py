"""Thin tensor-level wrappers over the C ABI (include/dcsnet.h).  Every function enqueues sm_100a kernels on the
current CUDA stream and returns immediately; nothing here computes on the host or falls back to PyTorch ops.

Activation tensors are channels-last complex: a float32 / float16 / bfloat16 tensor of shape (B, H, W, C, 2).

Every wrapper runs on the device of its tensor arguments (`_on_tensor_device`): the kernels are launched on that
device's current stream whatever `torch.cuda.current_device()` is.
"""
import ctypes as C
import functools

import torch

from . import _lib as L
from ._lib import F32, BF16, F16, ACT_NONE, ACT_RELU, ACT_LRELU  # noqa: F401

H16 = (torch.float16, torch.bfloat16)      # 16-bit activation storage types of the tensor-core modes

N_FFT, HOP, BINS = 512, 32, 256


def _code(t):
    return L.dtype_code(t)


def stft(audio, spec=None, bn_affine=None, bn_out=None, bn_real=False):
    """(B, L) fp32 -> (B, 256, T) complex64 [data.py:112-134]; optionally also the folded initial BN output (bn_real: of
    the magnitude, for the real path)."""
    L.require_cuda(audio)
    assert audio.dtype == torch.float32 and audio.dim() == 2 and audio.is_contiguous()
    B, n = audio.shape
    T = n // HOP + 1
    if spec is None:
        spec = torch.empty(B, BINS, T, dtype=torch.complex64, device=audio.device)
    p = L.StftParams(L.ptr(audio), L.ptr(spec), B, n, T, L.ptr(bn_affine), L.ptr(bn_out),
                     _code(bn_out) if bn_out is not None else F32, int(bn_real))
    L.check(L.lib().dcs_stft_fwd(C.byref(p), L.stream_ptr()), "dcs_stft_fwd")
    return spec


def istft(spec, audio=None, atan2_eps=1e-6, exact_polar=False):
    """(B, 256, T) complex64 -> (B, 32 (T-1)) fp32 [network_functions.py:398-401 + 140-150]."""
    L.require_cuda(spec)
    assert spec.dtype == torch.complex64 and spec.dim() == 3 and spec.shape[1] == BINS
    spec = spec.contiguous()
    B, _, T = spec.shape
    if audio is None:
        audio = torch.empty(B, HOP * (T - 1), dtype=torch.float32, device=spec.device)
    p = L.IstftParams(L.ptr(spec), L.ptr(audio), B, T, float(atan2_eps), int(exact_polar), None, None)
    L.check(L.lib().dcs_istft_fwd(C.byref(p), L.stream_ptr()), "dcs_istft_fwd")
    return audio


def istft_mag_phase(mag, phase, audio=None):
    """mag_phase_2_wave (network_functions.py:140-150) on explicit fp32 (B,256,T) magnitude / phase arrays."""
    L.require_cuda(mag, phase)
    assert mag.dtype == torch.float32 and phase.dtype == torch.float32 and mag.shape == phase.shape
    assert mag.dim() == 3 and mag.shape[1] == BINS and mag.is_contiguous() and phase.is_contiguous()
    B, _, T = mag.shape
    if audio is None:
        audio = torch.empty(B, HOP * (T - 1), dtype=torch.float32, device=mag.device)
    p = L.IstftParams(None, L.ptr(audio), B, T, 0.0, 1, L.ptr(mag), L.ptr(phase))
    L.check(L.lib().dcs_istft_fwd(C.byref(p), L.stream_ptr()), "dcs_istft_fwd")
    return audio


def cbn_apply(x, affine, y=None, act=ACT_NONE, out_dtype=None):
    """Folded eval-mode ComplexBatchNorm2d (+activation) on a channels-last complex tensor (..., C, 2)."""
    L.require_cuda(x, affine)
    Cn = x.shape[-2]
    if y is None:
        y = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    p = L.CbnParams(L.ptr(x), L.ptr(y), L.ptr(affine), x.numel() // (2 * Cn), Cn, act, _code(x), _code(y))
    L.check(L.lib().dcs_cbn_apply(C.byref(p), L.stream_ptr()), "dcs_cbn_apply")
    return y


def cconv(pk, src0, src1, dst, use_tc=False, pool_sums=None, pool_max=False):
    """Complex convolution with packed operands `pk` (packing.PackedConv).  src*: (B,H,W,C,2), dst: (B,OH,OW,Cout,2)."""
    L.require_cuda(src0, dst)
    B, H, W, c0, _ = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    assert c0 + c1 == pk.cin, (c0, c1, pk.cin)
    _, OH, OW, co, _ = dst.shape
    assert co == pk.cout
    p = L.CconvParams()
    p.src0, p.src1, p.c0, p.c1 = L.ptr(src0), L.ptr(src1), c0, c1
    p.batch, p.in_h, p.in_w = B, H, W
    p.out_h, p.out_w, p.cout = OH, OW, co
    p.up_h, p.up_w = pk.up
    p.stride_h, p.stride_w = pk.stride
    p.ntaps = pk.ntaps
    for i, (a, b) in enumerate(zip(pk.dy, pk.dx)):
        p.dy[i], p.dx[i] = a, b
    if use_tc:
        w = pk.w_tc if src0.dtype in H16 else pk.w_tc32
        assert w is not None, "PackedConv was not packed for this tensor-core operand type"
        assert src0.dtype not in H16 or w.dtype == src0.dtype, (w.dtype, src0.dtype)
    else:
        w = pk.w_ffma
    p.weight = L.ptr(w)
    p.bias = L.ptr(pk.bias)
    p.act = pk.act
    p.dst, p.in_dtype, p.out_dtype = L.ptr(dst), _code(src0), _code(dst)
    p.pool_sums, p.pool_mode = L.ptr(pool_sums), (L.POOL_MAX if pool_max else L.POOL_SUM)
    fn = L.lib().dcs_cconv2d_tc_fwd if use_tc else L.lib().dcs_cconv2d_fwd
    L.check(fn(C.byref(p), L.stream_ptr()), "dcs_cconv2d_tc_fwd" if use_tc else "dcs_cconv2d_fwd")
    return dst


def cconv_strip(sp, src0, src1, dst, pool_sums=None, tail=None, out_hw=None, pool_max=False):
    """Row-strip tensor-core convolution (fp16 / bf16) with operands `sp` (packing.StripConv / StripEnc0 / StripDec6).
    Same tensors as cconv(); with `tail` (a _lib.StripTail, decoder[6] only) dst is None and out_hw = (OH, OW)."""
    L.require_cuda(src0)
    pk = sp.pk
    B, H, W, c0, _ = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    assert (c0, c1) == (sp.c0, sp.c1) and src0.dtype == sp.w_image.dtype and src0.dtype in H16
    if tail is None:
        assert dst.dtype == src0.dtype
        _, OH, OW, co, _ = dst.shape
        assert co == pk.cout
    else:
        (OH, OW), co = out_hw, pk.cout
    p = L.CstripParams()
    p.src0, p.src1, p.c0, p.c1 = L.ptr(src0), L.ptr(src1), c0, c1
    p.batch, p.in_h, p.in_w = B, H, W
    p.out_h, p.out_w, p.cout = OH, OW, co
    p.up_h, p.up_w = sp.up
    p.stride_h, p.stride_w = sp.stride
    p.n_groups = len(sp.groups)
    for i, g in enumerate(sp.groups):
        for k, v in g.items():
            setattr(p.group[i], k, v)
    p.items, p.n_items_total = L.ptr(sp.item_table), sp.item_table.shape[0]
    p.weights = L.ptr(sp.w_image)
    p.box_units, p.n_mma, p.cols = sp.box_units, sp.n_mma, sp.cols
    p.bias, p.act = L.ptr(pk.bias), pk.act
    p.dst, p.pool_sums = L.ptr(dst), L.ptr(pool_sums)
    p.tail = C.pointer(tail) if tail is not None else None
    p.dtype, p.pool_mode = _code(src0), (L.POOL_MAX if pool_max else L.POOL_SUM)
    L.check(L.lib().dcs_cconv2d_strip_fwd(C.byref(p), L.stream_ptr()), "dcs_cconv2d_strip_fwd")
    return dst


def dec6_tail_strip(sp, d, skip, noisy_spec, clean_spec, net_raw=None, net_out=None, mask=None, noise_spec=None,
                    atan2_eps=1e-6, combine=L.COMBINE_DCS, exact_polar=False):
    """decoder[6] + bound_cRM x2 + combine on the tensor cores (row-strip kernel with the tail epilogue).
    sp = packing.StripDec6; d, skip (B, h, w, 8, 2) bf16; spectrogram arrays (B, 2h, 2w) complex64."""
    pk = sp.pk
    B, H, W, Cn, _ = d.shape
    assert Cn == 8 and skip.shape == d.shape and W % 4 == 0
    tail = L.StripTail(L.ptr(noisy_spec), L.ptr(net_raw), L.ptr(net_out), L.ptr(mask), L.ptr(noise_spec), L.ptr(clean_spec),
                       float(pk.bias_host[0]), float(pk.bias_host[1]), float(atan2_eps), combine, int(exact_polar))
    cconv_strip(sp, sp.view_src(d), sp.view_src(skip), None, tail=tail, out_hw=(2 * H, 2 * W))
    return clean_spec


def conv_out_hw(pk, H, W):
    if pk.up != (1, 1):
        return H * pk.up[0], W * pk.up[1]
    return (H + 2 * (pk.kh // 2) - pk.kh) // pk.stride[0] + 1, (W + 2 * (pk.kw // 2) - pk.kw) // pk.stride[1] + 1


def zero_(t):
    """Clear a device buffer with cudaMemsetAsync on the current stream (a memset node under graph capture)."""
    L.require_cuda(t)
    assert t.is_contiguous()
    L.check(L.lib().dcs_zero(L.ptr(t), t.numel() * t.element_size(), L.stream_ptr()), "dcs_zero")
    return t


def pool_sums_to_float(sums):
    """int64 fixed-point pooled sums (include/dcsnet.h: DCS_POOL_FRAC_BITS) -> float64 tensor (host-side inspection)."""
    return sums.double() / float(1 << L.POOL_FRAC_BITS)


def chan_pool(x, sums):
    """sums (B, C, 2) int64 fixed point, pre-zeroed (zero_())."""
    assert sums.dtype == torch.int64
    B, H, W, Cn, _ = x.shape
    p = L.ChanPoolParams(L.ptr(x), L.ptr(sums), B, H * W, Cn, _code(x))
    L.check(L.lib().dcs_chan_pool(C.byref(p), L.stream_ptr()), "dcs_chan_pool")


def pool_mean(sums, hw, mean=None):
    """int64 fixed-point pooled sums (B, C, 2) -> fp32 means (B, C, 2) = sums / hw."""
    L.require_cuda(sums)
    assert sums.dtype == torch.int64 and sums.is_contiguous()
    if mean is None:
        mean = torch.empty(sums.shape, dtype=torch.float32, device=sums.device)
    L.check(L.lib().dcs_pool_mean(L.ptr(sums), 1.0 / hw, L.ptr(mean), sums.numel(), L.stream_ptr()), "dcs_pool_mean")
    return mean


def chan_gate(sums, hw, ca, gate):
    B = sums.shape[0]
    p = L.ChanGateParams(L.ptr(sums), 1.0 / hw, L.ptr(gate), B, ca["channels"], ca["reduced"],
                         L.ptr(ca["w1_r"]), L.ptr(ca["w1_i"]), L.ptr(ca["w2_r"]), L.ptr(ca["w2_i"]))
    L.check(L.lib().dcs_chan_gate(C.byref(p), L.stream_ptr()), "dcs_chan_gate")


def spat_stats(x, gate, stats, sums=None, ca=None, gate_out=None):
    """Per-pixel channel statistics of gate * x.  With sums + ca (pack_channel_attention) the channel-gate MLP runs
    inside the kernel (no dcs_chan_gate launch) and the gate is written to gate_out for dcs_spat_apply."""
    B, H, W, Cn, _ = x.shape
    if sums is not None:
        p = L.SpatStatsParams(L.ptr(x), None, L.ptr(stats), B, H, W, Cn, _code(x), L.ptr(sums), ca["reduced"],
                              L.ptr(ca["w1_r"]), L.ptr(ca["w1_i"]), L.ptr(ca["w2_r"]), L.ptr(ca["w2_i"]), L.ptr(gate_out))
    else:
        p = L.SpatStatsParams(L.ptr(x), L.ptr(gate), L.ptr(stats), B, H, W, Cn, _code(x), None, 0, None, None, None, None, None)
    L.check(L.lib().dcs_spat_stats(C.byref(p), L.stream_ptr()), "dcs_spat_stats")


def spat_apply(x, gate, stats, w7, y, gate_out=None):
    B, H, W, Cn, _ = x.shape
    p = L.SpatApplyParams(L.ptr(x), L.ptr(gate), L.ptr(stats), L.ptr(w7), L.ptr(y), B, H, W, Cn, _code(x),
                          _code(y) if y is not None else F32, L.ptr(gate_out))
    L.check(L.lib().dcs_spat_apply(C.byref(p), L.stream_ptr()), "dcs_spat_apply")


def attention_fused(x, sums, ca, w7, y):
    """y = SA(CA(x) * x) * (CA(x) * x) in one pass; sums (B,C,2) = sum over H*W of x; ca = pack_channel_attention()."""
    L.require_cuda(x, sums, y)
    B, H, W, Cn, _ = x.shape
    p = L.AttentionParams(L.ptr(x), L.ptr(y), L.ptr(sums), B, H, W, Cn, ca["reduced"], _code(x), _code(y),
                          L.ptr(ca["w1_r"]), L.ptr(ca["w1_i"]), L.ptr(ca["w2_r"]), L.ptr(ca["w2_i"]), L.ptr(w7), 0)
    L.check(L.lib().dcs_attention_fused(C.byref(p), L.stream_ptr()), "dcs_attention_fused")
    return y


def attention_stream(x, sums, ca, w7, y):
    """attention_fused() as the streaming row-ring kernel (fp16 / bf16 storage; TF32 tensor-core gate conv)."""
    L.require_cuda(x, sums, y)
    B, H, W, Cn, _ = x.shape
    p = L.AttentionParams(L.ptr(x), L.ptr(y), L.ptr(sums), B, H, W, Cn, ca["reduced"], _code(x), _code(y),
                          L.ptr(ca["w1_r"]), L.ptr(ca["w1_i"]), L.ptr(ca["w2_r"]), L.ptr(ca["w2_i"]), L.ptr(w7), 0)
    L.check(L.lib().dcs_attention_stream(C.byref(p), L.stream_ptr()), "dcs_attention_stream")
    return y


def real_attention_stream(x, maxima, att, w7, y):
    """Real CBAM (r_network.py:8-40) as the streaming row-ring kernel on a PAIR tensor x (B, H, W, C/2, 2) of 2 * (C/2) real
    channels, fp16 storage.  maxima (B, C/2, 2) int64: the per-(image, channel) maxima in the conv epilogues' DCS_POOL_MAX
    encoding; att = dict(w1 (R, C), w2 (C, R)); w7 = conv1.weight flattened (2, 49)."""
    L.require_cuda(x, maxima, y)
    B, H, W, Cp, _ = x.shape
    assert maxima.dtype == torch.int64 and att["w1"].shape[1] == 2 * Cp
    p = L.AttentionParams(L.ptr(x), L.ptr(y), L.ptr(maxima), B, H, W, Cp, att["w1"].shape[0], _code(x), _code(y),
                          L.ptr(att["w1"]), None, L.ptr(att["w2"]), None, L.ptr(w7), 1)
    L.check(L.lib().dcs_attention_stream(C.byref(p), L.stream_ptr()), "dcs_attention_stream")
    return y


def chan_max(x, maxima):
    """Per-(image, real channel) maxima of a pair tensor into pre-zeroed int64 `maxima` (B, C/2, 2), DCS_POOL_MAX encoding
    (stand-alone form of the conv epilogues' pool_max; used when the producing kernel could not fuse it)."""
    L.require_cuda(x, maxima)
    B, H, W, Cp, _ = x.shape
    p = L.ChanPoolParams(L.ptr(x), L.ptr(maxima), B, H * W, Cp, _code(x))
    L.check(L.lib().dcs_chan_max(C.byref(p), L.stream_ptr()), "dcs_chan_max")


def real_attention(x, att, w7, y=None, workspace=None):
    """RealChannelAttention + RealSpatialAttention (r_network.py:8-40) on a channels-last real tensor (B, H, W, C) — or the
    same memory viewed as channel pairs (B, H, W, C/2, 2).  att = dict(w1 (R, C), w2 (C, R)), w7 = conv1.weight flattened."""
    L.require_cuda(x)
    shp = x.shape
    xr = x.reshape(shp[0], shp[1], shp[2], -1)
    B, H, W, Cn = xr.shape
    if y is None:
        y = torch.empty_like(x)
    need = int(L.lib().dcs_real_attention_workspace_bytes(B, H, W, Cn))
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
    p = L.RealAttentionParams(L.ptr(x), L.ptr(y), B, H, W, Cn, att["w1"].shape[0], _code(x), L.ptr(att["w1"]), L.ptr(att["w2"]),
                              L.ptr(w7), L.ptr(workspace), workspace.numel() * workspace.element_size())
    L.check(L.lib().dcs_real_attention_fwd(C.byref(p), L.stream_ptr()), "dcs_real_attention_fwd")
    return y


def rlstm(x, w, y=None, workspace=None):
    """nn.LSTM(D -> 128, 2 layers, bidirectional) of the real path: x (B, S, D) fp32 -> (B, S, 256) fp32.
    w = packing.PackedRNet(...).lstm_t."""
    L.require_cuda(x)
    B, S, D = x.shape
    H = w["w_hh_t"].shape[2]
    if y is None:
        y = torch.empty(B, S, 2 * H, dtype=torch.float32, device=x.device)
    need = int(L.lib().dcs_rlstm_workspace_bytes(B, S, H))
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
    p = L.RlstmParams(L.ptr(x), L.ptr(y), B, S, D, H, _code(x), L.ptr(w["w_ih0_t"]), L.ptr(w["w_ih1_t"]), L.ptr(w["w_hh_t"]),
                      L.ptr(w["bias"]), L.ptr(workspace), workspace.numel() * workspace.element_size())
    L.check(L.lib().dcs_rlstm_fwd(C.byref(p), L.stream_ptr()), "dcs_rlstm_fwd")
    return y


def rlstm_tc_workspace_bytes(B, S, hidden=128):
    n = L.lib().dcs_rlstm_tc_workspace_bytes(B, S, hidden)
    if n < 0:
        raise RuntimeError("dcs_rlstm_tc_workspace_bytes: unsupported shape")
    return int(n)


def rlstm_tc(x, w, y, workspace):
    """Tensor-core form of rlstm(): x (B, S, D) and y (B, S, 256) in the same 16-bit type; w = PackedRNet(...).lstm_tc."""
    L.require_cuda(x, y)
    B, S, D = x.shape
    assert x.dtype in H16 and y.dtype == x.dtype and w["w_ih0"].dtype == x.dtype and x.is_contiguous() and y.is_contiguous()
    p = L.RlstmTcParams(L.ptr(x), L.ptr(y), B, S, D, y.shape[2] // 2, _code(x), L.ptr(w["w_ih0"]), L.ptr(w["w_ih1"]), L.ptr(w["w_hh"]),
                        L.ptr(w["bias"]), L.ptr(workspace), workspace.numel() * workspace.element_size())
    L.check(L.lib().dcs_rlstm_tc_fwd(C.byref(p), L.stream_ptr()), "dcs_rlstm_tc_fwd")
    return y


def clstm_workspace_bytes(B, S, hidden=64):
    n = L.lib().dcs_clstm_workspace_bytes(B, S, hidden)
    if n < 0:
        raise RuntimeError("dcs_clstm_workspace_bytes: unsupported shape")
    return int(n)


def clstm(x, y, w, workspace, use_tc=False, seqs_per_cta=0):
    """x (B,S,D,2) -> y (B,S,2*hidden,2) fp32.  w = packing.pack_lstm(...).  use_tc: tf32 tensor-core projections."""
    B, S, D, _ = x.shape
    p = L.ClstmParams(L.ptr(x), L.ptr(y), B, S, D, y.shape[2] // 2, _code(x), L.ptr(w["w_ih0"]), L.ptr(w["w_ih1"]),
                      L.ptr(w["w_hh"]), L.ptr(w["bias"]), L.ptr(workspace), workspace.numel() * workspace.element_size(),
                      L.ptr(w["w_ih0_t"]) if use_tc else None, L.ptr(w["w_ih1_t"]) if use_tc else None, seqs_per_cta,
                      L.ptr(w["w_hh_frag"]) if use_tc else None)
    L.check(L.lib().dcs_clstm_fwd(C.byref(p), L.stream_ptr()), "dcs_clstm_fwd")
    return y


def mask_combine(net_raw, noisy_spec, clean_spec, net_out=None, mask=None, noise_spec=None, atan2_eps=1e-6,
                 combine=L.COMBINE_DCS, exact_polar=False):
    n = noisy_spec.numel()
    p = L.MaskCombineParams(L.ptr(net_raw), L.ptr(noisy_spec), L.ptr(net_out), L.ptr(mask), L.ptr(noise_spec),
                            L.ptr(clean_spec), n, float(atan2_eps), combine, int(exact_polar))
    L.check(L.lib().dcs_mask_combine(C.byref(p), L.stream_ptr()), "dcs_mask_combine")
    return clean_spec


def enc0(pk, spec, bn_affine, dst):
    """initial_batchnorm + encoder[0] (+BN+ReLU) straight from the complex64 spectrogram (B,F,T)."""
    B, F, T = spec.shape
    assert (pk.cin, pk.cout, pk.kh, pk.kw, pk.stride) == (1, 8, 7, 7, (2, 2)) and pk.act == ACT_RELU
    p = L.Enc0Params(L.ptr(spec), L.ptr(bn_affine), L.ptr(pk.w_ffma), L.ptr(pk.bias), L.ptr(dst), _code(dst), B, F, T)
    L.check(L.lib().dcs_enc0_fwd(C.byref(p), L.stream_ptr()), "dcs_enc0_fwd")
    return dst


def dec6_tail(pk, d, skip, noisy_spec, clean_spec, net_raw=None, net_out=None, mask=None, noise_spec=None,
              atan2_eps=1e-6, combine=L.COMBINE_DCS, exact_polar=False):
    """decoder[6] + bound_cRM x2 + combine in one kernel.  pk: PackedConv of decoder[6] (has .w_tail, .bias)."""
    B, H, W, Cn, _ = d.shape
    assert Cn == 8 and skip.shape == d.shape and pk.cout == 1 and pk.up == (2, 2)
    p = L.Dec6TailParams(L.ptr(d), L.ptr(skip), _code(d), B, H, W, L.ptr(pk.w_tail), float(pk.bias_host[0]),
                         float(pk.bias_host[1]), L.ptr(noisy_spec), L.ptr(net_raw), L.ptr(net_out), L.ptr(mask),
                         L.ptr(noise_spec), L.ptr(clean_spec), float(atan2_eps), combine, int(exact_polar))
    L.check(L.lib().dcs_dec6_tail_fwd(C.byref(p), L.stream_ptr()), "dcs_dec6_tail_fwd")
    return clean_spec


def convert(src, dst):
    L.check(L.lib().dcs_convert(L.ptr(src), L.ptr(dst), src.numel(), _code(src), _code(dst), L.stream_ptr()), "dcs_convert")
    return dst


def bound_crm(x, atan2_eps=1e-6, exact_polar=True):
    """bound_cRM (network_functions.py:77-88) on a complex64 tensor."""
    L.require_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    L.check(L.lib().dcs_bound_crm(L.ptr(x), L.ptr(y), x.numel(), float(atan2_eps), int(exact_polar), L.stream_ptr()), "dcs_bound_crm")
    return y


def cmul(a, b):
    L.require_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    assert a.shape == b.shape and a.dtype == torch.complex64 and b.dtype == torch.complex64
    y = torch.empty_like(a)
    L.check(L.lib().dcs_cmul(L.ptr(a), L.ptr(b), L.ptr(y), a.numel(), L.stream_ptr()), "dcs_cmul")
    return y


def mag_phase(spec, atan2_eps=10e-7, want_phase=True):
    """|spec| and atan2(im, re + eps) of a complex64 tensor (network_functions.py:286-288)."""
    L.require_cuda(spec)
    spec = spec.contiguous()
    mag = torch.empty(spec.shape, dtype=torch.float32, device=spec.device)
    phase = torch.empty_like(mag) if want_phase else None
    L.check(L.lib().dcs_mag_phase(L.ptr(spec), L.ptr(mag), L.ptr(phase), spec.numel(), float(atan2_eps), L.stream_ptr()), "dcs_mag_phase")
    return mag, phase


def real_mask_combine(mag, mask, subtract):
    """dr (subtract=False): clean = mag * mask; drs (True): noise = mag * mask, clean = mag - noise (network_functions.py:296-305,
    338-342).  `mask` may be the strided (B, F, T) view R_NETWORK.forward returns."""
    L.require_cuda(mag, mask)
    assert mag.is_contiguous() and mask.shape == mag.shape and mask.dtype == torch.float32
    st = mask.stride()
    exp, acc = [], 1
    for n_ in reversed(mask.shape):
        exp.append(acc); acc *= n_
    unit = st[-1]
    assert tuple(st) == tuple(unit * e for e in reversed(exp)), "mask must be a contiguous tensor or a uniformly strided view of one"
    clean = torch.empty_like(mag)
    noise = torch.empty_like(mag) if subtract else None
    L.check(L.lib().dcs_real_mask_combine(L.ptr(mag), L.ptr(mask), unit, L.ptr(clean), L.ptr(noise), mag.numel(), int(subtract),
                                          L.stream_ptr()), "dcs_real_mask_combine")
    return clean, noise


def crm(S, Y, eps=1e-8):
    L.require_cuda(S, Y)
    S, Y = S.contiguous(), Y.contiguous()
    m = torch.empty_like(S)
    L.check(L.lib().dcs_crm(L.ptr(S), L.ptr(Y), L.ptr(m), S.numel(), float(eps), L.stream_ptr()), "dcs_crm")
    return m


def upsample_nearest(x, up):
    B, H, W, Cn, _ = x.shape
    y = torch.empty(B, H * up[0], W * up[1], Cn, 2, dtype=x.dtype, device=x.device)
    L.check(L.lib().dcs_upsample_nearest(L.ptr(x), L.ptr(y), B, H, W, Cn, up[0], up[1], _code(x), L.stream_ptr()), "dcs_upsample_nearest")
    return y


# ---------------------------------------------------------------- device guard
def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its first CUDA-tensor argument current: `L.stream_ptr()` and the library's
    launches then target the tensors' device even when the caller's current device is another GPU."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


for _name, _fn in list(globals().items()):
    if callable(_fn) and getattr(_fn, "__module__", None) == __name__ and not _name.startswith("_") \
            and _name not in ("conv_out_hw", "clstm_workspace_bytes", "rlstm_tc_workspace_bytes", "pool_sums_to_float"):
        globals()[_name] = _on_tensor_device(_fn)
del _name, _fn
