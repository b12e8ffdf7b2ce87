"""ctypes binding of libpdse.so (the C ABI declared in include/pdse.h).

There is no CPU fallback: if the library is missing or the device is not sm_100 the
product path raises.  PyTorch only supplies device memory (``tensor.data_ptr()``) and
streams; no torch type crosses the ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

_LIB = None

_P = C.c_void_p
_I = C.c_int
_L = C.c_long
_F = C.c_float
_U64 = C.c_ulonglong

_SIGNATURES = {
    "pdse_abi_version": ([], _I),
    "pdse_operand_format": ([], _I),
    "pdse_check_device": ([], _I),
    "pdse_sm_count": ([], _I),
    "pdse_signal_table_floats": ([], _I),
    "pdse_signal_tables": ([_P], _I),
    "pdse_rms_f32": ([_P, _I, _I, _P, _P], _I),
    "pdse_rms_ragged_f32": ([_P, _P, _I, _I, _P, _P], _I),
    "pdse_stft_compress_ragged_f32": ([_P, _P, _P, _P, _P, _I, _I, _I, _P], _I),
    "pdse_decompress_istft_ragged_f32": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "pdse_decompress_istft_pcm16_f32": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "pdse_absmax_ragged_f32": ([_P, _P, _I, _I, _P, _P], _I),
    "pdse_stft_compress_f32": ([_P, _P, _P, _P, _I, _I, _I, _P], _I),
    "pdse_decompress_istft_f32": ([_P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "pdse_absmax_f32": ([_P, _I, _I, _P, _P], _I),
    "pdse_init_state_f32": ([_P, _P, _P, _L, _I, _I, _U64, _U64, _P], _I),
    "pdse_init_state_add_f32": ([_P, _P, _P, _P, _L, _I, _I, _U64, _U64, _P], _I),
    "pdse_ddpm_update_f32": ([_P, _P, _P, _P, _P, _L, _I, _F, _F, _F, _I, _I, _F, _U64, _U64, _P], _I),
    "pdse_scale_f32": ([_P, _L, _F, _P], _I),
    "pdse_randn_aten_policy": ([_L, _P, _P], _I),
    "pdse_randn_aten_f32": ([_P, _L, _U64, _U64, _I, _P], _I),
    "pdse_f32_to_pcm16": ([_P, _P, _L, _I, _P], _I),
    "pdse_ssnr_f32": ([_P, _P, _P, _I, _I, _P, _P], _I),
    "pdse_bias_row_floats": ([], _I),
    "pdse_time_embed": ([_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "pdse_enc1_fwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _P], _I),
    "pdse_enc_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "pdse_tcm_fwd": ([_P] * 12 + [_I, _I, _I, _P], _I),
    "pdse_tcm_flow": ([_P] * 12 + [_I, _I, _P], _I),
    "pdse_status_check": ([_P], _I),
    "pdse_dec_fwd": ([_P] * 11 + [_I] * 9 + [_P, _P], _I),
    "pdse_gcrn_conv1_fwd": ([_P, _P, _P, _P, _P, _I, _I, _P], _I),
    "pdse_gcrn_enc_fwd": ([_P] * 7 + [_I] * 6 + [_P], _I),
    "pdse_gcrn_dec_fwd": ([_P] * 6 + [_I] * 7 + [_P], _I),
    "pdse_lstm_inproj": ([_P, _P, _P, _P, _I, _I, _I, _P], _I),
    "pdse_lstm_rec": ([_P] * 8 + [_I, _I, _I, _P], _I),
    "pdse_debug_lstm_prof": ([_P], _I),
    "pdse_debug_lstm_clusters": ([_I], _I),
    "pdse_debug_dec_prof": ([_P], _I),
    "pdse_debug_tcm_prof": ([_P], _I),
    "pdse_gcrn_ln": ([_P] * 7 + [_I, _I, _I, _P], _I),
    "pdse_gcrn_out_fwd": ([_P] * 6 + [_I, _I, _P], _I),
    "pdse_db_guard_frames": ([], _I),
    "pdse_db_conv_fwd": ([_P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P], _I),
    "pdse_db_ln_fwd": ([_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P], _I),
    "pdse_aia_attn_fwd": ([_P, _P, _P, _P, _I, _I, _I, _P], _I),
    "pdse_aia_gru_fwd": ([_P, _P, _P, _P, _I, _I, _P], _I),
    "pdse_aia_post_fwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _P], _I),
    "pdse_aia_combine_fwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P], _I),
    "pdse_aia_aham_fwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _P], _I),
    "pdse_dw_guard_rows": ([], _I),
    "pdse_dw_embed": ([_P, _I] + [_P] * 7 + [_I, _P, _P], _I),
    "pdse_dw_pre_fwd": ([_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _P], _I),
    "pdse_dw_layer_fwd": ([_P] * 7 + [_I] * 6 + [_P], _I),
    "pdse_dw_post_fwd": ([_P, _P, _P, _F, _P, _I, _I, _P], _I),
    "pdse_pack_layout": ([_I, _P, _I, _P], _L),
    "pdse_pack_diffunet1": ([_P, _I, _P], _I),
    "pdse_pack_gcrn": ([_P, _I, _P], _I),
    "pdse_workspace_bytes": ([_I, _I, _I], _L),
    "pdse_diffunet1_time_bias": ([_P, _P, _I, _P, _P], _I),
    "pdse_diffunet1_fwd": ([_P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _P], _I),
    "pdse_gcrn_fwd": ([_P, _P, _P, _P, _I, _I, _P], _I),
    "pdse_probe_tmem": ([_P, _P, _I, _I, _I, _I, _I, _I, _L, _I, _P], _I),
    "pdse_probe_gemm": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
}


def exported_symbols():
    return ["pdse_last_error"] + list(_SIGNATURES)


def load(require_device: bool = False):
    """dlopen the in-tree library (building it first if sources are newer)."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if not os.path.exists(path) or _build.stale():
            try:
                path = _build.build()
            except Exception as e:  # no nvcc on the box: use what shipped
                if not os.path.exists(path):
                    raise RuntimeError(f"libpdse.so is missing and could not be built: {e}") from e
        lib = C.CDLL(path)
        lib.pdse_last_error.argtypes = []
        lib.pdse_last_error.restype = C.c_char_p
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        _LIB = lib
    if require_device:
        if not torch.cuda.is_available():
            raise RuntimeError("prior_diffuse_b200 needs a CUDA device (sm_100a); there is no CPU path")
        check(_LIB.pdse_check_device())
    return _LIB


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libpdse: " + load().pdse_last_error().decode())


def op_dtype():
    """torch dtype of the library's 16-bit tensor-core operand format (csrc/opfmt.h): float16 unless built with bf16 operands"""
    return torch.float16 if load().pdse_operand_format() == 1 else torch.bfloat16


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


# ---------------------------------------------------------------------------- host-side packing through the C ABI
NET_DIFFUNET1, NET_GCRN = 1, 2


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_long)]


class BlobEntry(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("dtype", C.c_int), ("offset", C.c_long), ("elems", C.c_long)]


def pack_layout(net: int):
    """directory of a packed network: ({name: (dtype, byte offset, elements)}, blob bytes)"""
    L = load()
    n = C.c_int(0)
    L.pdse_pack_layout(net, None, 0, C.byref(n))
    arr = (BlobEntry * n.value)()
    total = L.pdse_pack_layout(net, arr, n.value, C.byref(n))
    if total < 0:
        raise RuntimeError("libpdse: " + L.pdse_last_error().decode())
    return {e.name.decode(): (e.dtype, e.offset, e.elems) for e in arr}, total


def pack_state_dict(net: int, state_dict):
    """reference state_dict (torch tensors) -> (uint8 host blob, directory) through pdse_pack_*"""
    import numpy as np
    L = load()
    directory, total = pack_layout(net)
    keep, items = [], []
    for k, v in state_dict.items():
        if not v.is_floating_point():
            continue
        a = np.ascontiguousarray(v.detach().cpu().float().numpy())
        keep.append(a)
        items.append(Tensor(k.encode(), a.ctypes.data, a.size))
    arr = (Tensor * len(items))(*items)
    blob = np.zeros(total, dtype=np.uint8)
    fn = L.pdse_pack_diffunet1 if net == NET_DIFFUNET1 else L.pdse_pack_gcrn
    check(fn(arr, len(items), blob.ctypes.data_as(C.c_void_p)))
    return blob, directory
