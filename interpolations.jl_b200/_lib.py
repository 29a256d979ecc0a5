"""ctypes binding of libb200itp.so (the C ABI declared in include/b200itp.h).

There is no CPU fallback: if the CUDA library is missing the import of this module
raises, and every product entry point goes through it.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_DIMS = 4
MAX_WOODBURY = 8

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200ITP_LIB selects another build of the same library (A/B timing of build-time variants, tools/ only)
LIB_PATH = os.environ.get("B200ITP_LIB") or os.path.join(_HERE, "libb200itp.so")


class Desc(C.Structure):
    """b200itp_desc (include/b200itp.h)."""
    _fields_ = [
        ("ndims", C.c_int32),
        ("coef_dtype", C.c_int32),
        ("coord_dtype", C.c_int32),
        ("etp_pair_mask", C.c_int32),
        ("size", C.c_int64 * MAX_DIMS),
        ("parent_len", C.c_int64 * MAX_DIMS),
        ("coef_first", C.c_int32 * MAX_DIMS),
        ("degree", C.c_int32 * MAX_DIMS),
        ("bc", C.c_int32 * MAX_DIMS),
        ("gridtype", C.c_int32 * MAX_DIMS),
        ("inner_lb", C.c_double * MAX_DIMS),
        ("inner_ub", C.c_double * MAX_DIMS),
        ("inner_tag", C.c_int32 * MAX_DIMS),
        ("has_scale", C.c_int32),
        ("range_tag", C.c_int32 * MAX_DIMS),
        ("range_first", C.c_double * MAX_DIMS),
        ("range_step", C.c_double * MAX_DIMS),
        ("outer_lb", C.c_double * MAX_DIMS),
        ("outer_ub", C.c_double * MAX_DIMS),
        ("outer_tag", C.c_int32 * MAX_DIMS),
        ("etp_kind", C.c_int32),
        ("etp_lo", C.c_int32 * MAX_DIMS),
        ("etp_hi", C.c_int32 * MAX_DIMS),
        ("fill_tag", C.c_int32),
        ("fill_value", C.c_double),
        ("parent_first", C.c_int32 * MAX_DIMS),
    ]


class AxisSystem(C.Structure):
    """b200itp_axis_system (include/b200itp.h)."""
    _fields_ = [
        ("dtype", C.c_int32),
        ("k", C.c_int32),
        ("n", C.c_int64),
        ("dl", C.c_void_p),
        ("d", C.c_void_p),
        ("du", C.c_void_p),
        ("urow", C.POINTER(C.c_int32)),
        ("vcol", C.POINTER(C.c_int32)),
        ("Cp", C.c_void_p),
    ]


class B200ItpError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libb200itp.so; raise loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200ItpError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the product path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    pi32, pi64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sigs = {
        "b200itp_version": (C.c_int, []),
        "b200itp_last_error": (sz, [C.c_char_p, sz]),
        "b200itp_launch_count": (C.c_uint64, []),
        "b200itp_profile_enable": (C.c_int, [i32]),
        "b200itp_profile_count": (C.c_int, []),
        "b200itp_profile_get": (C.c_int, [i32, C.c_char_p, sz, C.POINTER(C.c_double)]),
        "b200itp_malloc": (C.c_int, [i32, sz, C.POINTER(vp)]),
        "b200itp_free": (C.c_int, [i32, vp]),
        "b200itp_memcpy_h2d": (C.c_int, [i32, vp, vp, sz, vp]),
        "b200itp_memcpy_d2h": (C.c_int, [i32, vp, vp, sz, vp]),
        "b200itp_sync": (C.c_int, [i32, vp]),
        "b200itp_copy_with_padding": (C.c_int, [i32, vp, vp, i32, i32, pi64, pi32, vp]),
        "b200itp_prefiltering_system": (C.c_int, [i32, i64, i32, i32, i32, vp, vp, vp, pi32, pi32, pi32, vp]),
        "b200itp_prefilter_axis": (C.c_int, [i32, vp, i32, i32, pi64, i32, C.POINTER(AxisSystem), vp]),
        "b200itp_prefilter": (C.c_int, [i32, vp, i32, i32, pi64, pi32, pi32, pi32, vp]),
        "b200itp_prefilter_from": (C.c_int, [i32, vp, vp, i32, i32, pi64, pi32, pi32, pi32, vp]),
        "b200itp_prefilter_padded_from": (C.c_int, [i32, vp, vp, i32, i32, pi64, pi32, pi32, pi32, pi32, vp]),
        "b200itp_eval": (C.c_int, [i32, C.POINTER(Desc), vp, C.POINTER(vp), i64, vp, vp, pi64, vp, vp, vp]),
        "b200itp_eval_grid_scratch_bytes": (sz, [C.POINTER(Desc), pi64]),
        "b200itp_eval_grid": (C.c_int, [i32, C.POINTER(Desc), vp, C.POINTER(vp), pi64, vp, pi64, vp, sz, vp]),
        "b200itp_hessian": (C.c_int, [i32, C.POINTER(Desc), vp, C.POINTER(vp), i64, vp, pi64, vp]),
        "b200itp_out_dtype": (C.c_int, [C.POINTER(Desc)]),
        "b200itp_eval_sorted_scratch_bytes": (sz, [C.POINTER(Desc), i64]),
        "b200itp_eval_sorted": (C.c_int, [i32, C.POINTER(Desc), vp, C.POINTER(vp), i64, vp, pi64, vp, sz, vp]),
        "b200itp_gradient_sorted": (C.c_int, [i32, C.POINTER(Desc), vp, C.POINTER(vp), i64, vp, pi64, vp, sz, vp]),
        "b200itp_eval_sorted_min_queries": (i64, []),
        "b200itp_eval_sorted_set_min_queries": (None, [i64]),
        "b200itp_eval_sorted_set_tma": (C.c_int, [C.c_int]),
        "b200itp_eval_sorted_tma_calls": (C.c_uint64, []),
        "b200itp_prefilter_set_segmentation": (C.c_int, [C.c_int]),
        "b200itp_prefilter_set_strict": (C.c_int, [C.c_int]),
        "b200itp_eval_grid_set_separable": (C.c_int, [C.c_int]),
        "b200itp_eval_set_coefs_event": (C.c_int, [vp]),
        "b200itp_set_l2_fetch_granularity": (C.c_int, [i32, i32]),
        "b200itp_comm_init": (C.c_int, [i32, pi32, C.POINTER(vp)]),
        "b200itp_comm_size": (C.c_int, [vp]),
        "b200itp_comm_destroy": (C.c_int, [vp]),
        "b200itp_bcast_coefs": (C.c_int, [vp, i32, C.POINTER(vp), sz]),
        "b200itp_prefilter_sharded": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), i32, pi64, pi32, pi32, pi32,
                                                C.POINTER(C.c_double)]),
        "b200itp_eval_grid_separable_calls": (C.c_uint64, []),
        "b200itp_eval_sorted_ordered_calls": (C.c_uint64, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "b200itp_version", "b200itp_last_error", "b200itp_launch_count", "b200itp_profile_enable",
    "b200itp_profile_count", "b200itp_profile_get", "b200itp_malloc", "b200itp_free",
    "b200itp_memcpy_h2d", "b200itp_memcpy_d2h", "b200itp_sync", "b200itp_copy_with_padding",
    "b200itp_prefiltering_system", "b200itp_prefilter_axis", "b200itp_prefilter", "b200itp_prefilter_from", "b200itp_prefilter_padded_from", "b200itp_eval", "b200itp_eval_grid", "b200itp_eval_grid_scratch_bytes", "b200itp_hessian",
    "b200itp_out_dtype", "b200itp_eval_sorted_scratch_bytes", "b200itp_eval_sorted", "b200itp_gradient_sorted",
    "b200itp_eval_sorted_min_queries", "b200itp_eval_sorted_set_min_queries", "b200itp_eval_sorted_set_tma", "b200itp_eval_sorted_tma_calls", "b200itp_prefilter_set_segmentation",
    "b200itp_prefilter_set_strict", "b200itp_eval_sorted_ordered_calls", "b200itp_eval_grid_set_separable",
    "b200itp_eval_grid_separable_calls", "b200itp_eval_set_coefs_event", "b200itp_set_l2_fetch_granularity", "b200itp_comm_init", "b200itp_comm_size", "b200itp_comm_destroy",
    "b200itp_bcast_coefs", "b200itp_prefilter_sharded",
]


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().b200itp_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str):
    """0 ok; >0 unsupported/invalid (ArgumentError in the Julia glue); <0 CUDA error."""
    if rc == 0:
        return
    msg = last_error()
    if rc > 0:
        raise ValueError(f"{what}: unsupported or invalid argument (code {rc}): {msg}")
    raise B200ItpError(f"{what}: CUDA error (code {rc}): {msg}")


def profile_stages():
    """Per-stage device times recorded since b200itp_profile_enable(1): {name: [ms, ...]} in launch order."""
    L = lib()
    out = {}
    buf = C.create_string_buffer(96)
    ms = C.c_double(0.0)
    for i in range(L.b200itp_profile_count()):
        check(L.b200itp_profile_get(i, buf, 96, C.byref(ms)), "b200itp_profile_get")
        out.setdefault(buf.value.decode(), []).append(ms.value)
    return out
