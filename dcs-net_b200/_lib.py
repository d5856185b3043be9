"""ctypes binding of libdcsnet_sm100a.so (include/dcsnet.h).  No torch types cross this boundary: only raw
device pointers (`tensor.data_ptr()`), sizes and the current CUDA stream handle.

The library is mandatory: there is NO CPU / PyTorch fallback.  `lib()` raises if the shared object is missing or
does not export every symbol declared in include/dcsnet.h.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcsnet_sm100a.so")

F32, BF16, F16 = 0, 1, 2
POOL_SUM, POOL_MAX = 0, 1
POOL_FRAC_BITS = 28           # pooled sums are int64 fixed point (include/dcsnet.h: DCS_POOL_FRAC_BITS)
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3
COMBINE_DCS, COMBINE_DC, COMBINE_DR, COMBINE_DRS = 0, 1, 2, 3
MAX_TAPS = 64

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64


class StftParams(C.Structure):
    _fields_ = [("audio", _vp), ("spec", _vp), ("batch", _i), ("length", _i), ("n_frames", _i),
                ("bn_affine", _vp), ("bn_out", _vp), ("bn_dtype", _i), ("bn_real", _i)]


class FrontendParams(C.Structure):
    _fields_ = [("clean48", _vp), ("noisy48", _vp), ("lengths48", _vp), ("start16", _vp), ("batch", _i), ("stride48", _i64),
                ("window", _i), ("kernel", _vp), ("n_taps", _i), ("orig", _i), ("width", _i),
                ("clean16", _vp), ("noisy16", _vp), ("noise16", _vp), ("flags", _vp)]


class RealAttentionParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("batch", _i), ("h", _i), ("w", _i), ("channels", _i), ("reduced", _i), ("dtype", _i),
                ("w1", _vp), ("w2", _vp), ("w7", _vp), ("workspace", _vp), ("workspace_bytes", _i64)]


class RlstmParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("batch", _i), ("seq", _i), ("in_dim", _i), ("hidden", _i), ("in_dtype", _i),
                ("w_ih0_t", _vp), ("w_ih1_t", _vp), ("w_hh_t", _vp), ("bias", _vp), ("workspace", _vp), ("workspace_bytes", _i64)]


class RlstmTcParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("batch", _i), ("seq", _i), ("in_dim", _i), ("hidden", _i), ("dtype", _i),
                ("w_ih0", _vp), ("w_ih1", _vp), ("w_hh", _vp), ("bias", _vp), ("workspace", _vp), ("workspace_bytes", _i64)]


class CbnTrainParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("n_pix", _i64), ("channels", _i), ("act", _i), ("in_dtype", _i), ("out_dtype", _i),
                ("weight", _vp), ("bias", _vp), ("eps", _f), ("momentum", _f),
                ("running_mean", _vp), ("running_covar", _vp), ("num_batches_tracked", _vp),
                ("affine", _vp), ("saved", _vp), ("workspace", _vp), ("workspace_bytes", _i64)]


class CbnTrainBwdParams(C.Structure):
    _fields_ = [("x", _vp), ("dy", _vp), ("dx", _vp), ("n_pix", _i64), ("channels", _i),
                ("saved", _vp), ("weight", _vp), ("dweight", _vp), ("dbias", _vp), ("workspace", _vp), ("workspace_bytes", _i64),
                ("conv_bias_grad_r", _vp), ("conv_bias_grad_i", _vp), ("x_dtype", _i)]


class CwgradParams(C.Structure):
    _fields_ = [("x", _vp), ("dy", _vp), ("dtype", _i),
                ("batch", _i), ("in_h", _i), ("in_w", _i), ("out_h", _i), ("out_w", _i), ("cin", _i), ("cout", _i),
                ("stride_h", _i), ("stride_w", _i), ("ntaps", _i), ("dy_off", C.c_int8 * MAX_TAPS), ("dx_off", C.c_int8 * MAX_TAPS),
                ("dw_r", _vp), ("dw_i", _vp), ("workspace", _vp), ("workspace_bytes", _i64)]


class WgradParams(C.Structure):
    _fields_ = [("x", _vp), ("dy", _vp),
                ("batch", _i), ("in_h", _i), ("in_w", _i), ("out_h", _i), ("out_w", _i), ("k2", _i), ("n2", _i), ("x_pitch", _i),
                ("dy_pitch", _i), ("stride_h", _i), ("stride_w", _i), ("ntaps", _i), ("dy_off", C.c_int8 * MAX_TAPS),
                ("dx_off", C.c_int8 * MAX_TAPS), ("dwp", _vp), ("workspace", _vp), ("workspace_bytes", _i64)]


class Wgrad16Params(C.Structure):
    _fields_ = [("x", _vp), ("dy", _vp), ("dtype", _i),
                ("batch", _i), ("in_h", _i), ("in_w", _i), ("out_h", _i), ("out_w", _i), ("k2", _i), ("n2", _i), ("x_pitch", _i),
                ("dy_pitch", _i), ("stride_h", _i), ("stride_w", _i), ("ntaps", _i), ("dy_off", C.c_int8 * MAX_TAPS),
                ("dx_off", C.c_int8 * MAX_TAPS), ("dwp", _vp), ("dwp_tap_stride", _i64), ("workspace", _vp), ("workspace_bytes", _i64)]


class AttentionBwdParams(C.Structure):
    _fields_ = [("x", _vp), ("dy", _vp), ("gate_c", _vp), ("stats", _vp), ("gate_s", _vp), ("w7", _vp), ("sums", _vp),
                ("batch", _i), ("h", _i), ("w", _i), ("channels", _i), ("reduced", _i),
                ("w1_r", _vp), ("w1_i", _vp), ("w2_r", _vp), ("w2_i", _vp),
                ("dspre", _vp), ("dx", _vp), ("chan_const", _vp),
                ("dw1_r", _vp), ("dw1_i", _vp), ("dw2_r", _vp), ("dw2_i", _vp), ("dw7_r", _vp), ("dw7_i", _vp),
                ("workspace", _vp), ("workspace_bytes", _i64), ("x_dtype", _i)]


class IstftParams(C.Structure):
    _fields_ = [("spec", _vp), ("audio", _vp), ("batch", _i), ("n_frames", _i), ("atan2_eps", _f), ("exact_polar", _i),
                ("mag", _vp), ("phase", _vp)]


class CbnParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("affine", _vp), ("n_pix", _i64), ("channels", _i), ("act", _i),
                ("in_dtype", _i), ("out_dtype", _i)]


class CconvParams(C.Structure):
    _fields_ = [("src0", _vp), ("src1", _vp), ("c0", _i), ("c1", _i),
                ("batch", _i), ("in_h", _i), ("in_w", _i),
                ("out_h", _i), ("out_w", _i), ("cout", _i),
                ("up_h", _i), ("up_w", _i), ("stride_h", _i), ("stride_w", _i),
                ("ntaps", _i), ("dy", C.c_int8 * MAX_TAPS), ("dx", C.c_int8 * MAX_TAPS),
                ("weight", _vp), ("bias", _vp), ("act", _i),
                ("dst", _vp), ("in_dtype", _i), ("out_dtype", _i),
                ("pool_sums", _vp), ("pool_mode", _i), ("bias_phase_stride", _i)]


STRIP_MAX_GROUPS = 2


class StripGroup(C.Structure):
    _fields_ = [("item0", _i), ("n_items", _i), ("dy_min", _i), ("n_dy", _i), ("ph0", _i), ("n_ph", _i), ("x_min", _i),
                ("w_bytes", _i), ("w_off", _i64)]


class StripTail(C.Structure):
    _fields_ = [("noisy_spec", _vp), ("net_raw", _vp), ("net_out", _vp), ("mask", _vp), ("noise_spec", _vp),
                ("clean_spec", _vp), ("bias_re", _f), ("bias_im", _f), ("atan2_eps", _f), ("combine", _i), ("exact_polar", _i)]


class CstripParams(C.Structure):
    _fields_ = [("src0", _vp), ("src1", _vp), ("c0", _i), ("c1", _i),
                ("batch", _i), ("in_h", _i), ("in_w", _i),
                ("out_h", _i), ("out_w", _i), ("cout", _i),
                ("up_h", _i), ("up_w", _i), ("stride_h", _i), ("stride_w", _i),
                ("n_groups", _i), ("group", StripGroup * STRIP_MAX_GROUPS),
                ("items", _vp), ("n_items_total", _i),
                ("weights", _vp),
                ("box_units", _i), ("n_mma", _i), ("cols", _i),
                ("bias", _vp), ("act", _i),
                ("dst", _vp), ("pool_sums", _vp), ("tail", C.POINTER(StripTail)), ("dtype", _i), ("pool_mode", _i)]


class ChanPoolParams(C.Structure):
    _fields_ = [("x", _vp), ("sums", _vp), ("batch", _i), ("hw", _i), ("channels", _i), ("dtype", _i)]


class ChanGateParams(C.Structure):
    _fields_ = [("sums", _vp), ("inv_hw", _f), ("gate", _vp), ("batch", _i), ("channels", _i), ("reduced", _i),
                ("w1_r", _vp), ("w1_i", _vp), ("w2_r", _vp), ("w2_i", _vp)]


class SpatStatsParams(C.Structure):
    _fields_ = [("x", _vp), ("chan_gate", _vp), ("stats", _vp), ("batch", _i), ("h", _i), ("w", _i),
                ("channels", _i), ("dtype", _i),
                ("sums", _vp), ("reduced", _i), ("w1_r", _vp), ("w1_i", _vp), ("w2_r", _vp), ("w2_i", _vp), ("gate_out", _vp)]


class SpatApplyParams(C.Structure):
    _fields_ = [("x", _vp), ("chan_gate", _vp), ("stats", _vp), ("w7", _vp), ("y", _vp),
                ("batch", _i), ("h", _i), ("w", _i), ("channels", _i), ("in_dtype", _i), ("out_dtype", _i),
                ("gate_out", _vp)]


class AttentionParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("sums", _vp), ("batch", _i), ("h", _i), ("w", _i), ("channels", _i),
                ("reduced", _i), ("in_dtype", _i), ("out_dtype", _i),
                ("w1_r", _vp), ("w1_i", _vp), ("w2_r", _vp), ("w2_i", _vp), ("w7", _vp), ("real", _i)]


class ClstmParams(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("batch", _i), ("seq", _i), ("in_dim", _i), ("hidden", _i), ("in_dtype", _i),
                ("w_ih0", _vp), ("w_ih1", _vp), ("w_hh", _vp), ("bias", _vp),
                ("workspace", _vp), ("workspace_bytes", _i64), ("w_ih0_t", _vp), ("w_ih1_t", _vp),
                ("seqs_per_cta", _i), ("w_hh_frag", _vp)]


class MaskCombineParams(C.Structure):
    _fields_ = [("net_raw", _vp), ("noisy_spec", _vp), ("net_out", _vp), ("mask", _vp), ("noise_spec", _vp),
                ("clean_spec", _vp), ("n", _i64), ("atan2_eps", _f), ("combine", _i), ("exact_polar", _i)]


class Enc0Params(C.Structure):
    _fields_ = [("spec", _vp), ("bn_affine", _vp), ("weight", _vp), ("bias", _vp), ("dst", _vp), ("out_dtype", _i),
                ("batch", _i), ("h", _i), ("w", _i)]


class Dec6TailParams(C.Structure):
    _fields_ = [("d", _vp), ("skip", _vp), ("in_dtype", _i), ("batch", _i), ("h", _i), ("w", _i),
                ("weight", _vp), ("bias_re", _f), ("bias_im", _f),
                ("noisy_spec", _vp), ("net_raw", _vp), ("net_out", _vp), ("mask", _vp), ("noise_spec", _vp),
                ("clean_spec", _vp), ("atan2_eps", _f), ("combine", _i), ("exact_polar", _i)]


# symbol -> (restype, argtypes); must list EVERY entry of include/dcsnet.h (tests/test_abi.py checks both ways)
SYMBOLS = {
    "dcs_abi_version": (_i, []),
    "dcs_last_error_string": (C.c_char_p, []),
    "dcs_launch_count": (C.c_uint64, []),
    "dcs_real_attention_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "dcs_real_attention_fwd": (_i, [C.POINTER(RealAttentionParams), _vp]),
    "dcs_rlstm_workspace_bytes": (_i64, [_i, _i, _i]),
    "dcs_rlstm_fwd": (_i, [C.POINTER(RlstmParams), _vp]),
    "dcs_rlstm_tc_workspace_bytes": (_i64, [_i, _i, _i]),
    "dcs_rlstm_tc_fwd": (_i, [C.POINTER(RlstmTcParams), _vp]),
    "dcs_cbn_train_workspace_bytes": (_i64, [_i64, _i]),
    "dcs_cbn_train_fwd": (_i, [C.POINTER(CbnTrainParams), _vp]),
    "dcs_cbn_train_bwd": (_i, [C.POINTER(CbnTrainBwdParams), _vp]),
    "dcs_si_snr": (_i, [_vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp]),
    "dcs_istft_adjoint": (_i, [_vp, _vp, _i, _i, _vp]),
    "dcs_mask_tail_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _vp]),
    "dcs_upcat_adjoint": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dcs_cwgrad_workspace_bytes": (_i64, [C.POINTER(CwgradParams)]),
    "dcs_cwgrad_tc": (_i, [C.POINTER(CwgradParams), _vp]),
    "dcs_wgrad_workspace_bytes": (_i64, [C.POINTER(WgradParams)]),
    "dcs_wgrad": (_i, [C.POINTER(WgradParams), _vp]),
    "dcs_wgrad_fold_complex": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dcs_transpose": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dcs_sgemm": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dcs_colsum_workspace_bytes": (_i64, [_i64, _i]),
    "dcs_colsum": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _i64, _vp]),
    "dcs_dilate": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dcs_cconv_dgrad_cin1": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dcs_upcat_fwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dcs_wgrad_tc16_workspace_bytes": (_i64, [C.POINTER(Wgrad16Params)]),
    "dcs_wgrad_tc16": (_i, [C.POINTER(Wgrad16Params), _vp]),
    "dcs_dec6_bwd_workspace_bytes": (_i64, []),
    "dcs_dec6_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "dcs_act_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _vp]),
    "dcs_dropout": (_i, [_vp, _vp, _i64, _i, _f, C.c_uint64, C.c_uint64, _vp]),
    "dcs_attention_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "dcs_attention_bwd": (_i, [C.POINTER(AttentionBwdParams), _vp]),
    "dcs_lstm_train_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dcs_lstm_train_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dcs_cplx_split": (_i, [_vp, _i, _vp, _i64, _vp]),
    "dcs_cplx_merge": (_i, [_vp, _vp, _i64, _vp]),
    "dcs_clstm_combine": (_i, [_vp, _vp, _i64, _vp]),
    "dcs_clstm_combine_bwd": (_i, [_vp, _vp, _i64, _vp]),
    "dcs_sumsq": (_i, [_vp, _i64, _vp, _i, _vp, _i64, _vp]),
    "dcs_adam_amsgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _vp, _f, _f, _vp]),
    "dcs_gather_pack": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "dcs_frontend_fwd": (_i, [C.POINTER(FrontendParams), _vp]),
    "dcs_stft_fwd": (_i, [C.POINTER(StftParams), _vp]),
    "dcs_istft_fwd": (_i, [C.POINTER(IstftParams), _vp]),
    "dcs_cbn_apply": (_i, [C.POINTER(CbnParams), _vp]),
    "dcs_cconv2d_fwd": (_i, [C.POINTER(CconvParams), _vp]),
    "dcs_cconv2d_tc_fwd": (_i, [C.POINTER(CconvParams), _vp]),
    "dcs_cconv2d_strip_fwd": (_i, [C.POINTER(CstripParams), _vp]),
    "dcs_chan_pool": (_i, [C.POINTER(ChanPoolParams), _vp]),
    "dcs_chan_max": (_i, [C.POINTER(ChanPoolParams), _vp]),
    "dcs_chan_gate": (_i, [C.POINTER(ChanGateParams), _vp]),
    "dcs_pool_mean": (_i, [_vp, _f, _vp, _i64, _vp]),
    "dcs_spat_stats": (_i, [C.POINTER(SpatStatsParams), _vp]),
    "dcs_spat_apply": (_i, [C.POINTER(SpatApplyParams), _vp]),
    "dcs_attention_fused": (_i, [C.POINTER(AttentionParams), _vp]),
    "dcs_attention_stream": (_i, [C.POINTER(AttentionParams), _vp]),
    "dcs_clstm_workspace_bytes": (_i64, [_i, _i, _i]),
    "dcs_clstm_fwd": (_i, [C.POINTER(ClstmParams), _vp]),
    "dcs_mask_combine": (_i, [C.POINTER(MaskCombineParams), _vp]),
    "dcs_dec6_tail_fwd": (_i, [C.POINTER(Dec6TailParams), _vp]),
    "dcs_enc0_fwd": (_i, [C.POINTER(Enc0Params), _vp]),
    "dcs_bound_crm": (_i, [_vp, _vp, _i64, _f, _i, _vp]),
    "dcs_cmul": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "dcs_crm": (_i, [_vp, _vp, _vp, _i64, _f, _vp]),
    "dcs_mag_phase": (_i, [_vp, _vp, _vp, _i64, _f, _vp]),
    "dcs_real_mask_combine": (_i, [_vp, _vp, _i64, _vp, _vp, _i64, _i, _vp]),
    "dcs_upsample_nearest": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dcs_tc_set_debug_buffer": (_i, [_vp]),
    "dcs_convert": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "dcs_zero": (_i, [_vp, _i64, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL.  Raises RuntimeError when the native library is unavailable."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python dcs-net_b200/build.py` "
                "(or __graft_entry__.build()).  dcsnet_b200 has no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            try:
                fn = getattr(l, name)
            except AttributeError as e:
                raise RuntimeError(f"{LIB_PATH} does not export {name}; rebuild it") from e
            fn.restype, fn.argtypes = res, args
        if l.dcs_abi_version() != 2:
            raise RuntimeError("libdcsnet_sm100a.so ABI version mismatch; rebuild it")
        _lib = l
    return _lib


def launch_count():
    return int(lib().dcs_launch_count())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        msg = lib().dcs_last_error_string()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dcsnet_b200 runs on CUDA tensors only (sm_100a kernels; no CPU fallback)")


def dtype_code(t):
    if t.dtype in (torch.float32, torch.complex64):
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise RuntimeError(f"unsupported activation dtype {t.dtype}")
