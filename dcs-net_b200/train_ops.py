"""Tensor-level wrappers of the TRAINING-step kernels (include/dcsnet.h, section f2; csrc/train.cu, the ADJ mode of csrc/stft.cu)
and the host-side packing of the convolution dgrad.  Same conventions as ops.py: channels-last complex tensors (..., C, 2),
everything enqueued on the current stream of the tensors' device, no PyTorch-op fallback.

Reference: network_functions.py:210-280 (train_batch_2_loss), 168-208 (calc_loss), 30-42 (SiSNR); c_network.py:243-261;
complexPyTorch 0.3 ComplexBatchNorm2d in train mode.  Oracles: oracle/train_oracle.py.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops, packing

BN_MOMENTUM = 0.1


def cbn_train_workspace(n_pix, channels, device):
    n = int(L.lib().dcs_cbn_train_workspace_bytes(n_pix, channels))
    if n < 0:
        raise RuntimeError("dcs_cbn_train_workspace_bytes: unsupported shape (channels must be a power of two <= 256)")
    return torch.empty(n, dtype=torch.uint8, device=device)


@ops._on_tensor_device
def cbn_train_fwd(x, weight, bias, running_mean=None, running_covar=None, num_batches_tracked=None, act=L.ACT_NONE, y=None,
                  eps=packing.BN_EPS, momentum=BN_MOMENTUM, want_saved=True, workspace=None):
    """Train-mode ComplexBatchNorm2d (+ activation) on x (..., C, 2): batch statistics, running-stat update IN PLACE
    (running_mean: complex64 (C,), running_covar (C, 3), num_batches_tracked int64 ()).  Returns (y, saved (C, 8), affine (C, 6))."""
    L.require_cuda(x, weight, bias)
    Cn = x.shape[-2]
    n_pix = x.numel() // (2 * Cn)
    assert x.is_contiguous() and weight.dtype == torch.float32 and tuple(weight.shape) == (Cn, 3) and tuple(bias.shape) == (Cn, 2)
    if y is None:
        y = torch.empty_like(x)
    ws = workspace if workspace is not None else cbn_train_workspace(n_pix, Cn, x.device)
    affine = torch.empty(Cn, 6, dtype=torch.float32, device=x.device)
    saved = torch.empty(Cn, 8, dtype=torch.float32, device=x.device) if want_saved else None
    rm = None
    if running_mean is not None:
        assert running_mean.dtype == torch.complex64 and running_mean.is_contiguous() and running_covar.is_contiguous()
        rm = torch.view_as_real(running_mean)
    p = L.CbnTrainParams(L.ptr(x), L.ptr(y), n_pix, Cn, act, L.dtype_code(x), L.dtype_code(y), L.ptr(weight.contiguous()),
                         L.ptr(bias.contiguous()), float(eps), float(momentum), L.ptr(rm), L.ptr(running_covar),
                         L.ptr(num_batches_tracked), L.ptr(affine), L.ptr(saved), L.ptr(ws), ws.numel())
    L.check(L.lib().dcs_cbn_train_fwd(C.byref(p), L.stream_ptr()), "dcs_cbn_train_fwd")
    return y, saved, affine


@ops._on_tensor_device
def cbn_train_bwd(x, dy, saved, weight, workspace=None):
    """Backward of cbn_train_fwd (no activation): x, dy fp32 (..., C, 2) -> (dx, dweight (C, 3), dbias (C, 2))."""
    L.require_cuda(x, dy, saved, weight)
    Cn = x.shape[-2]
    n_pix = x.numel() // (2 * Cn)
    assert x.dtype == torch.float32 and dy.dtype == torch.float32 and x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape
    dx = torch.empty_like(x)
    dw = torch.empty(Cn, 3, dtype=torch.float32, device=x.device)
    db = torch.empty(Cn, 2, dtype=torch.float32, device=x.device)
    ws = workspace if workspace is not None else cbn_train_workspace(n_pix, Cn, x.device)
    p = L.CbnTrainBwdParams(L.ptr(x), L.ptr(dy), L.ptr(dx), n_pix, Cn, L.ptr(saved), L.ptr(weight.contiguous()), L.ptr(dw), L.ptr(db),
                            L.ptr(ws), ws.numel())
    L.check(L.lib().dcs_cbn_train_bwd(C.byref(p), L.stream_ptr()), "dcs_cbn_train_bwd")
    return dx, dw, db


@ops._on_tensor_device
def si_snr(clean, estimate, grad_scale=None, eps=1e-8):
    """SiSNR(clean, estimate) per row (network_functions.py:30-42).  Returns (values (B,), grad (B, L) or None) with
    grad = grad_scale / B * d value / d estimate."""
    L.require_cuda(clean, estimate)
    assert clean.shape == estimate.shape and clean.dim() == 2 and clean.dtype == torch.float32 and estimate.dtype == torch.float32
    clean, estimate = clean.contiguous(), estimate.contiguous()
    B, n = clean.shape
    val = torch.empty(B, dtype=torch.float32, device=clean.device)
    grad = torch.empty_like(estimate) if grad_scale is not None else None
    L.check(L.lib().dcs_si_snr(L.ptr(clean), L.ptr(estimate), B, n, float(eps), float(grad_scale or 0.0), L.ptr(val), L.ptr(grad),
                               L.stream_ptr()), "dcs_si_snr")
    return val, grad


@ops._on_tensor_device
def istft_adjoint(grad_audio, n_frames):
    """Adjoint of mag_phase_2_wave's iSTFT: (B, 32 (T-1)) -> (B, 256, T) complex64 (dL/dRe + j dL/dIm)."""
    L.require_cuda(grad_audio)
    B, n = grad_audio.shape
    assert n == 32 * (n_frames - 1) and grad_audio.dtype == torch.float32
    g = torch.empty(B, 256, n_frames, dtype=torch.complex64, device=grad_audio.device)
    L.check(L.lib().dcs_istft_adjoint(L.ptr(grad_audio.contiguous()), L.ptr(g), B, n_frames, L.stream_ptr()), "dcs_istft_adjoint")
    return g


@ops._on_tensor_device
def mask_tail_bwd(net_raw, noisy_spec, g_clean, g_noise=None, atan2_eps=10e-7):
    """Adjoint of the fused mask tail: spectrogram-domain gradients of the clean (and noise, dcs) estimates -> d / d raw."""
    L.require_cuda(net_raw, noisy_spec, g_clean)
    for t in (net_raw, noisy_spec, g_clean, g_noise):
        assert t is None or (t.dtype == torch.complex64 and t.is_contiguous() and t.shape == net_raw.shape)
    d = torch.empty_like(net_raw)
    L.check(L.lib().dcs_mask_tail_bwd(L.ptr(net_raw), L.ptr(noisy_spec), L.ptr(g_clean), L.ptr(g_noise), L.ptr(d), net_raw.numel(),
                                      float(atan2_eps), L.stream_ptr()), "dcs_mask_tail_bwd")
    return d


@ops._on_tensor_device
def upcat_adjoint(g, c0, c1, up):
    """g (B, H*uh, W*uw, c0 + c1, 2) fp32 -> (gd (B, H, W, c0, 2), gskip (B, H, W, c1, 2)): the adjoint of cat + nearest up-sampling."""
    L.require_cuda(g)
    B, HH, WW, Cn, _ = g.shape
    uh, uw = up
    assert Cn == c0 + c1 and HH % uh == 0 and WW % uw == 0 and g.dtype == torch.float32 and g.is_contiguous()
    H, W = HH // uh, WW // uw
    gd = torch.empty(B, H, W, c0, 2, dtype=torch.float32, device=g.device)
    gs = torch.empty(B, H, W, c1, 2, dtype=torch.float32, device=g.device) if c1 else None
    L.check(L.lib().dcs_upcat_adjoint(L.ptr(g), L.ptr(gd), L.ptr(gs), B, H, W, c0, c1, uh, uw, L.stream_ptr()), "dcs_upcat_adjoint")
    return gd, gs


# ------------------------------------------------------------------------------------------------ conv dgrad = a conv
def dgrad_conv(w_r, w_i, transposed, device, tc_dtype=None, want_tf32=False):
    """The data gradient of a STRIDE-1 complex conv layer as the operands of the forward conv kernels (role-swapped weights):
      ComplexConvTranspose2d(k3, s1, p1) (decoder, c_network.py:135-147): its adjoint is the PLAIN conv with the same weight
          tensors, i.e. dX = ComplexConv2d(weight viewed as (Cout_t -> Cin_t)) applied to CONJUGATE-free real block form;
      ComplexConv2d(k, s1, p=k//2): its adjoint is the transposed conv = flipped, in/out-swapped conv.
    In the packed real formulation (oracle/train_oracle.cconv2d_backward) dX = conv_transpose(dY, Wp) with
    Wp = [[w_r, -w_i], [w_i, w_r]]: the block matrix is TRANSPOSED, which for the complex pair means w_i -> -w_i.
    Returns a packing.PackedConv (no bias, no activation) mapping dY (B, H, W, Cout, 2) -> dX (B, H, W, Cin, 2)."""
    w_r, w_i = w_r.detach().double().cpu(), w_i.detach().double().cpu()
    if transposed:
        # forward: y = convT(x, W) with W (Cin, Cout, k, k); adjoint: dx = conv(dy, W) (cross-correlation with W as (out=Cin, in=Cout))
        wr, wi = w_r, -w_i
    else:
        # forward: y = conv(x, W) with W (Cout, Cin, k, k); adjoint: dx = convT(dy, W) = conv(dy, flip(W^T))
        wr, wi = w_r.permute(1, 0, 2, 3).flip(2, 3), -w_i.permute(1, 0, 2, 3).flip(2, 3)
    return packing.PackedConv(wr, wi, None, None, device=device, tc_dtype=tc_dtype, want_tf32=want_tf32)


@ops._on_tensor_device
def cwgrad(x, dy, kernel, stride):
    """Weight gradient of ComplexConv2d(cin -> cout, kernel, stride, padding = kernel // 2) on the tensor cores: x (B, H, W, cin, 2),
    dy (B, OH, OW, cout, 2) in the same 16-bit storage type -> (dw_r, dw_i) fp32 (cout, cin, kh, kw) = conv_r / conv_i .weight.grad."""
    L.require_cuda(x, dy)
    kh, kw = (kernel, kernel) if isinstance(kernel, int) else kernel
    B, H, W, cin, _ = x.shape
    _, OH, OW, cout, _ = dy.shape
    assert x.dtype in ops.H16 and dy.dtype == x.dtype and x.is_contiguous() and dy.is_contiguous()
    p = L.CwgradParams()
    p.x, p.dy, p.dtype = L.ptr(x), L.ptr(dy), L.dtype_code(x)
    p.batch, p.in_h, p.in_w, p.out_h, p.out_w, p.cin, p.cout = B, H, W, OH, OW, cin, cout
    p.stride_h, p.stride_w = stride
    p.ntaps = kh * kw
    for ky in range(kh):
        for kx in range(kw):
            p.dy_off[ky * kw + kx], p.dx_off[ky * kw + kx] = ky - kh // 2, kx - kw // 2
    dw_r = torch.empty(cout, cin, kh, kw, dtype=torch.float32, device=x.device)
    dw_i = torch.empty_like(dw_r)
    n = int(L.lib().dcs_cwgrad_workspace_bytes(C.byref(p)))
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=x.device)
    p.dw_r, p.dw_i, p.workspace, p.workspace_bytes = L.ptr(dw_r), L.ptr(dw_i), L.ptr(ws), ws.numel()
    L.check(L.lib().dcs_cwgrad_tc(C.byref(p), L.stream_ptr()), "dcs_cwgrad_tc")
    return dw_r, dw_i
