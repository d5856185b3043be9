"""CPU: the C-ABI shared library loads and exports exactly the symbols include/dcsnet.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

import dcsnet_b200 as D
from dcsnet_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dcsnet.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(dcs_[a-z0-9_]+)\s*\(", txt))


def test_header_and_binding_agree():
    hs = header_symbols()
    assert hs == set(L.SYMBOLS), hs ^ set(L.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = L.lib()  # raises if the .so is missing or a symbol is absent: no fallback
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.dcs_abi_version() == 2


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert header_symbols() <= exported


def test_struct_sizes_match_header():
    """ctypes mirrors of the POD parameter structs must have the C compiler's layout."""
    src = '#include <stdio.h>\n#include "dcsnet.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(dcs_stft_params),sizeof(dcs_istft_params),sizeof(dcs_cbn_params),sizeof(dcs_cconv_params),' \
          'sizeof(dcs_chan_pool_params),sizeof(dcs_chan_gate_params),sizeof(dcs_spat_stats_params),' \
          'sizeof(dcs_spat_apply_params),sizeof(dcs_clstm_params),sizeof(dcs_mask_combine_params),' \
          'sizeof(dcs_strip_item),sizeof(dcs_strip_group),sizeof(dcs_strip_tail),sizeof(dcs_cstrip_params),' \
          'sizeof(dcs_attention_params),sizeof(dcs_enc0_params),sizeof(dcs_dec6_tail_params),sizeof(dcs_frontend_params),'\
          'sizeof(dcs_real_attention_params),sizeof(dcs_rlstm_params));return 0;}'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split()))
    mine = [ctypes.sizeof(t) for t in (L.StftParams, L.IstftParams, L.CbnParams, L.CconvParams, L.ChanPoolParams,
                                       L.ChanGateParams, L.SpatStatsParams, L.SpatApplyParams, L.ClstmParams,
                                       L.MaskCombineParams)] + [16] + \
           [ctypes.sizeof(t) for t in (L.StripGroup, L.StripTail, L.CstripParams, L.AttentionParams, L.Enc0Params, L.Dec6TailParams,
                                       L.FrontendParams, L.RealAttentionParams, L.RlstmParams)]
    assert sizes == mine


def test_argument_errors_are_reported_not_swallowed():
    """Validation happens before any CUDA call, so it is testable without a GPU."""
    lib = L.lib()
    p = L.StftParams()
    rc = lib.dcs_stft_fwd(ctypes.byref(p), None)
    assert rc != 0 and b"null pointer" in lib.dcs_last_error_string()
    q = L.CconvParams()
    assert lib.dcs_cconv2d_tc_fwd(ctypes.byref(q), None) != 0
    r = L.CstripParams()
    assert lib.dcs_cconv2d_strip_fwd(ctypes.byref(r), None) != 0 and b"null pointer" in lib.dcs_last_error_string()
    assert lib.dcs_clstm_workspace_bytes(2, 10, 32) < 0  # only hidden=64 is built
    assert lib.dcs_clstm_workspace_bytes(2, 10, 64) > 0
    # entries added for the next rows (front-end, real path): same contract
    for params, fn in ((L.FrontendParams, lib.dcs_frontend_fwd), (L.RealAttentionParams, lib.dcs_real_attention_fwd),
                       (L.RlstmParams, lib.dcs_rlstm_fwd), (L.AttentionParams, lib.dcs_attention_stream)):
        assert fn(ctypes.byref(params()), None) != 0 and b"null pointer" in lib.dcs_last_error_string()
    assert lib.dcs_rlstm_workspace_bytes(2, 10, 128) == 2 * 10 * 10 * 128 * 4 and lib.dcs_rlstm_workspace_bytes(0, 10, 128) < 0
    assert lib.dcs_real_attention_workspace_bytes(2, 4, 5, 16) > 0 and lib.dcs_real_attention_workspace_bytes(2, 0, 5, 16) < 0
    assert lib.dcs_mag_phase(None, None, None, 4, 1e-6, None) != 0 and b"bad arguments" in lib.dcs_last_error_string()


def test_training_struct_sizes_match_header():
    """Same check for the parameter structs of the training-step entries (f2)."""
    names = ["dcs_cbn_train_params", "dcs_cbn_train_bwd_params", "dcs_cwgrad_params", "dcs_wgrad_params", "dcs_wgrad16_params", "dcs_attention_bwd_params"]
    src = '#include <stdio.h>\n#include "dcsnet.h"\nint main(){printf("' + " ".join(["%zu"] * len(names)) + '\\n",' + \
          ",".join(f"sizeof({n})" for n in names) + ');return 0;}'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split()))
    mine = [ctypes.sizeof(t) for t in (L.CbnTrainParams, L.CbnTrainBwdParams, L.CwgradParams, L.WgradParams, L.Wgrad16Params, L.AttentionBwdParams)]
    assert sizes == mine


def test_training_entries_validate_arguments_without_a_gpu():
    lib = L.lib()
    assert lib.dcs_wgrad(ctypes.byref(L.WgradParams()), None) != 0 and b"null pointer" in lib.dcs_last_error_string()
    assert lib.dcs_wgrad_tc16(ctypes.byref(L.Wgrad16Params()), None) != 0 and b"null pointer" in lib.dcs_last_error_string()
    assert lib.dcs_attention_bwd(ctypes.byref(L.AttentionBwdParams()), None) != 0 and b"null pointer" in lib.dcs_last_error_string()
    assert lib.dcs_attention_bwd_workspace_bytes(2, 4, 5, 24, 2) < 0 and lib.dcs_attention_bwd_workspace_bytes(2, 4, 5, 32, 2) > 0
    assert lib.dcs_lstm_train_fwd(None, None, 4, 2, 10, 64, None, None, None, None) != 0 and b"bad arguments" in lib.dcs_last_error_string()
    assert lib.dcs_adam_amsgrad(None, None, None, None, None, 10, 1e-4, 0.9, 0.999, 1e-6, 0.0, 1, None, 100.0, 1.0, None) != 0
    assert lib.dcs_colsum_workspace_bytes(0, 4) < 0
