"""ctypes binding of libdic_b200.so (the C ABI declared in include/dic_b200.h).

The shared library is the product: there is NO CPU fallback.  If it has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``python -m
deep_interpolation_clustering_b200.build``) every operator raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdic_b200.so")

DIC_OK = 0
DIC_ERR_INVALID_ARGUMENT = -1
DIC_ERR_UNSUPPORTED = -2
DIC_ERR_CUDA = -3
DIC_ERR_NO_DEVICE = -4

_P = c_void_p  # every device pointer crosses the boundary as a plain address

# name -> (restype, argtypes); mirrors include/dic_b200.h one to one
SIGNATURES = {
    "dic_last_error": (c_char_p, []),
    "dic_version": (c_int, []),
    "dic_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "dic_sci_fwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int64, _P]),
    "dic_interp_bwd_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "dic_sci_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int64, _P]),
    "dic_cci_fwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P]),
    "dic_cci_bwd_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "dic_cci_sci_bwd_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "dic_cci_sci_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "dic_cci_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "dic_rbf_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int64, _P]),
    "dic_rbf_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int64, _P]),
    "dic_upload_encounters": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "dic_pack_encounters_host": (c_int64, [_P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_int64, POINTER(c_int)]),
    "dic_upload_encounters_packed": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "dic_expand_encounters": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "dic_dec_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dic_dec_q_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P]),
    "dic_dec_p": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "dic_dec_q_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P]),
    "dic_dec_kl_fwd_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float,
                                   c_float, _P]),
    "dic_kmeans_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dic_kmeans_assign": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "dic_kmeans_update": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "dic_kmeans_lloyd_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "dic_kmeans_lloyd_run": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int,
                                     c_double, _P]),
    "dic_kmeans_min_d2": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "dic_pairwise_dist_sum": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P]),
    "dic_pairwise_dist_sum_part": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "dic_pairwise_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "dic_cluster_rowsums": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "dic_cluster_scatter_workspace_bytes": (c_size_t, [c_int]),
    "dic_cluster_scatter": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "dic_dunn_minmax": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "dic_lstm_project_packed_bytes": (c_size_t, [c_int]),
    "dic_lstm_pack_wih": (c_int, [_P, _P, c_int, _P]),
    "dic_lstm_project": (c_int, [_P, c_int64, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "dic_lstm_packed_bytes": (c_size_t, []),
    "dic_lstm_pack_whh": (c_int, [_P, _P, _P, _P]),
    "dic_lstm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, _P]),
    "dic_lstm_bwd_step": (c_int, [_P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, _P]),
    "dic_colsum_workspace_bytes": (c_size_t, [c_int]),
    "dic_colsum_f32": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "dic_probe_mufu": (c_int, [POINTER(c_double), _P]),
    "dic_probe_ffma": (c_int, [POINTER(c_double), _P]),
}

_lib = None


class DicError(RuntimeError):
    """A libdic_b200 call returned a non-zero status."""


def lib():
    """The loaded shared library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -m deep_interpolation_clustering_b200.build`). "
                "There is no CPU fallback for the B200 hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so is stale: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().dic_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    """Raise the Python exception matching a dic_status (ValueError for argument /
    shape errors, as torch / sklearn would; RuntimeError for CUDA failures)."""
    if status == DIC_OK:
        return
    msg = f"{what}: {last_error()} (status {status})"
    if status in (DIC_ERR_INVALID_ARGUMENT, DIC_ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise DicError(msg)


def ptr(t):
    """Device address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
