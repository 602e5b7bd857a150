"""ctypes binding of libocn_b200.so -- the one and only compute backend of this package.

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCN_B200_LIB") or os.path.join(HERE, "libocn_b200.so")  # override: A/B of kernel variants
HEADER = os.path.join(HERE, "..", "include", "ocn_b200.h")

_P = c_void_p
_SIGS = {
    "ocn_abi_version": (c_int, []),
    "ocn_last_error": (ctypes.c_char_p, []),
    "ocn_device_sm_count": (c_int, []),
    "ocn_graph_validate": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "ocn_graph_build_bytes": (c_size_t, [c_int64, c_int]),
    "ocn_graph_build_count": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, _P, c_size_t, _P, _P, _P]),
    "ocn_graph_build_fill": (c_int, [_P, c_int64, c_int, c_int64, c_int64, _P, _P, _P]),
    "ocn_graph_mask_bytes": (c_size_t, [c_int64, c_int64]),
    "ocn_graph_mask_count": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, c_int64, c_int, _P, _P, c_size_t, _P, _P, _P]),
    "ocn_graph_mask_fill": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P]),
    "ocn_graph_select_count": (c_int, [_P, _P, c_int64, _P, _P]),
    "ocn_graph_select_fill": (c_int, [_P, _P, _P, _P, c_int64, c_float, _P, _P, _P, _P]),
    "ocn_rows_intersect_count": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, c_int64, _P, _P]),
    "ocn_rows_intersect_fill": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, c_int64, _P, _P, _P]),
    "ocn_rows_difference_count": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, c_int64, _P, _P]),
    "ocn_rows_difference_fill": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, c_int64, _P, _P, _P]),
    "ocn_rows_gather_count": (c_int, [_P, c_int64, _P, c_int64, _P, _P]),
    "ocn_rows_gather_fill": (c_int, [_P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, _P]),
    "ocn_rows_hadamard_count": (c_int, [_P, _P, _P, _P, c_int64, _P, _P]),
    "ocn_rows_hadamard_fill": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P]),
    "ocn_csr_colsum": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "ocn_cn_plan_bytes": (c_size_t, [c_int64]),
    "ocn_cn_colstat_bytes": (c_size_t, [c_int64]),
    "ocn_cn_record_bytes": (c_size_t, []),
    "ocn_cn_plan": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int64, _P, c_size_t, _P, _P]),
    "ocn_cn_build": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, _P, _P, _P, c_int64, _P,
                             c_int64, _P, _P, c_size_t, _P, _P]),
    "ocn_cn_hub_bytes": (c_size_t, [c_int64, c_int64, _P]),
    "ocn_cn_hub_timing_events": (c_int, [_P, _P]),
    "ocn_cn_hub_scratch_reset": (c_int, [_P]),
    "ocn_set_option": (c_int, [c_int, c_int64]),
    "ocn_get_option": (c_int64, [c_int]),
    "ocn_launch_count": (ctypes.c_longlong, []),
    "ocn_cn_stats": (c_int, [_P, _P, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, c_int,
                             _P, _P, _P, _P, _P, _P]),
    "ocn_cn_aggregate": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P,
                                 _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "ocn_cn_aggregate_bwd": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P,
                                     _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "ocn_cn_extract_count": (c_int, [_P, c_int64, _P, c_int64, c_int, c_int, _P, _P, _P, _P]),
    "ocn_cn_extract_fill": (c_int, [_P, _P, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P,
                                    _P, _P, _P, _P, _P, _P, _P, _P]),
    "ocn_cn_release": (c_int, [_P, _P, c_int64, _P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "ocn_spmm_csr": (c_int, [_P, _P, _P, c_int64, _P, c_int64, c_int, _P, _P]),
    "ocn_spmm_csr_bwd": (c_int, [_P, _P, _P, c_int64, _P, c_int64, c_int, _P, _P]),
    "ocn_spmm_csr_max_bwd": (c_int, [_P, _P, _P, c_int64, _P, _P, c_int64, _P, _P]),
    "ocn_gcn_norm": (c_int, [_P, _P, c_int64, _P, _P]),
    "ocn_gcn_spmm": (c_int, [_P, _P, _P, c_int64, _P, c_int, _P, c_int64, _P, _P]),
    "ocn_spgemm_scratch_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "ocn_spgemm_a2_symbolic": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P]),
    "ocn_spgemm_a2_numeric": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "ocn_cn_head_params": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "ocn_cn_head": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int64, _P, _P, _P]),
    "ocn_linear_tc_prep_floats": (c_int64, [c_int, c_int]),
    "ocn_linear_tc_prep": (c_int, [_P, c_int, c_int, _P, _P]),
    "ocn_linear_tc": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, c_float, c_int, _P, _P, c_int, _P, _P]),
    "ocn_mrr": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "ocn_hits_bytes": (c_size_t, [c_int64]),
    "ocn_hits_at_k": (c_int, [_P, c_int64, _P, c_int64, c_int64, _P, c_size_t, _P, _P]),
}

_lib = None


class OcnError(RuntimeError):
    pass


def header_symbols():
    """Every function name declared in include/ocn_b200.h."""
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocn_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OcnError(
                f"{LIB_PATH} is missing. Build it with `python -m ocn_b200.build` (needs nvcc). "
                "ocn_b200 has no CPU or eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.ocn_abi_version() != 3:
            raise OcnError("libocn_b200.so ABI version mismatch")
        _lib = L
    return _lib


OPTIONS = {"hub_window": 0, "hub_cta_window": 1, "hub_heavy_run": 2, "hub_walker": 3, "hub_seg_ctas": 4, "hub_exact": 5, "grouped_off": 6, "spgemm_mode": 7, "spmm_tma": 8, "head_tc": 9}


def set_option(name: str, value: int) -> None:
    """Tuning / test option of the library (include/ocn_b200.h OCN_OPT_*); 0 restores the default."""
    check(lib().ocn_set_option(OPTIONS[name], int(value)), "ocn_set_option")


def reset_options() -> None:
    for k in OPTIONS.values():
        lib().ocn_set_option(k, 0)


def check(status: int, what: str):
    if status != 0:
        msg = lib().ocn_last_error()
        raise OcnError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
