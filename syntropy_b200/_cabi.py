"""ctypes view of libksg_b200.so -- one Python callable per symbol of include/ksg_b200.h.

Loading never falls back to anything: if the library cannot be built or loaded the
import of the compute path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_c_i64 = C.c_int64
_c_vp = C.c_void_p


class PsiStep(C.Structure):
    _fields_ = [
        ("subspace", C.c_int32),
        ("count_offset", C.c_int32),
        ("sign", C.c_int32),
        ("centered", C.c_int32),
        ("divisor", C.c_double),
    ]


class Prepared(C.Structure):
    """ksg_prepared of include/ksg_b200.h (host struct of device pointers)."""
    _fields_ = [
        ("n", C.c_int64),
        ("dim", C.c_int32),
        ("dim_padded", C.c_int32),
        ("primary", C.c_int32),
        ("axes_mode", C.c_int32),
        ("ld", C.c_int64),
        ("n_padded", C.c_int64),
        ("perm", C.c_void_p),
        ("sorted_f64", C.c_void_p),
        ("shadow_f32", C.c_void_p),
        ("axis_vals", C.c_void_p),
        ("axis_pos", C.c_void_p),
        ("center", C.c_void_p),
        ("maxabs", C.c_void_p),
        ("tile_box", C.c_void_p),
        ("sub_box", C.c_void_p),
        ("flags", C.c_void_p),
        ("curve_keys", C.c_void_p),
        ("curve_table", C.c_void_p),
        ("axes_ready", C.c_void_p),
    ]


class Cross(C.Structure):
    """ksg_cross of include/ksg_b200.h."""
    _fields_ = [
        ("nq", C.c_int64),
        ("ld", C.c_int64),
        ("perm", C.c_void_p),
        ("sorted_f64", C.c_void_p),
        ("shadow_f32", C.c_void_p),
        ("home", C.c_void_p),
        ("maxabs", C.c_void_p),
    ]


_c_prep = C.POINTER(Prepared)
_c_cross = C.POINTER(Cross)

# name -> (restype, argtypes); mirrors include/ksg_b200.h one to one
SIGNATURES = {
    "ksg_version": (C.c_int, []),
    "ksg_last_error": (C.c_char_p, []),
    "ksg_launch_count": (C.c_ulonglong, []),
    "ksg_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "ksg_check_finite": (C.c_int, [_c_vp, _c_i64, _c_vp, _c_vp]),
    "ksg_psi_table": (C.c_int, [_c_vp, _c_i64, _c_vp]),
    "ksg_knn_radius_workspace": (C.c_size_t, [_c_i64, _c_i64, C.c_int, C.c_int]),
    "ksg_knn_radius": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, _c_vp, _c_i64, _c_i64, _c_i64, C.c_int,
                                 _c_vp, _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_count_subspaces_workspace": (C.c_size_t, [_c_i64, _c_i64, C.c_int, C.c_int]),
    "ksg_count_subspaces": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, _c_vp, _c_i64, _c_i64, _c_i64,
                                      C.POINTER(C.c_uint32), C.c_int, _c_vp, _c_i64, C.c_int, C.c_int32,
                                      _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_psi_epilogue": (C.c_int, [_c_vp, _c_i64, C.c_int, C.POINTER(PsiStep), C.c_int, C.c_double,
                                   C.c_double, C.c_int, C.c_double, _c_vp, _c_i64, _c_vp, _c_vp]),
    "ksg_entropy_epilogue": (C.c_int, [_c_vp, _c_i64, C.c_int, C.c_double, C.c_double, _c_vp, _c_vp]),
    "ksg_kl_epilogue": (C.c_int, [_c_vp, _c_vp, _c_i64, C.c_int, C.c_double, _c_vp, _c_vp]),
    "ksg_sum_workspace": (C.c_size_t, [_c_i64]),
    "ksg_sum_f64": (C.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_subspace_radius_from_neighbours": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, _c_i64, _c_i64, _c_vp,
                                                      C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int,
                                                      _c_vp, _c_vp]),
    "ksg_paircount_fixed_radius": (C.c_int, [_c_vp, _c_vp, _c_i64, C.c_int, C.c_double, _c_i64, _c_i64,
                                             _c_vp, _c_vp]),
    "ksg_prepare_workspace": (C.c_size_t, [_c_i64, C.c_int]),
    "ksg_prepare_join": (C.c_int, [_c_prep, _c_vp]),
    "ksg_prepare": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, C.c_int, _c_vp, C.c_size_t, _c_prep, _c_vp]),
    "ksg_prepare_rows": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, C.c_int, C.POINTER(C.c_void_p), _c_vp, C.c_size_t,
                                   _c_prep, _c_vp]),
    "ksg_fast_knn_workspace": (C.c_size_t, [_c_i64, _c_i64, C.c_int, C.c_int]),
    "ksg_fast_knn": (C.c_int, [_c_prep, _c_i64, _c_i64, _c_vp, C.c_int, C.c_int, _c_vp, _c_vp, _c_vp, _c_vp,
                               C.c_size_t, _c_vp]),
    "ksg_fast_knn_idx": (C.c_int, [_c_prep, _c_i64, _c_i64, _c_vp, C.c_int, C.c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                   C.c_size_t, _c_vp]),
    "ksg_fast_knn_workspace_margin": (C.c_size_t, [_c_i64, _c_i64, C.c_int, C.c_int, C.c_int]),
    "ksg_fast_knn_margin": (C.c_int, [_c_prep, _c_i64, _c_i64, _c_vp, C.c_int, C.c_int, C.c_int, _c_vp, _c_vp, _c_vp,
                                      _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_cross_prepare_workspace": (C.c_size_t, [_c_i64, _c_i64, C.c_int]),
    "ksg_cross_prepare": (C.c_int, [_c_prep, _c_vp, _c_i64, _c_i64, _c_vp, C.c_size_t, _c_cross, _c_vp]),
    "ksg_fast_knn_cross": (C.c_int, [_c_prep, _c_cross, _c_i64, _c_i64, C.c_int, _c_vp, _c_vp, _c_vp, _c_vp,
                                     C.c_size_t, _c_vp]),
    "ksg_fast_count_workspace": (C.c_size_t, [_c_i64, _c_i64, C.c_int, C.c_int]),
    "ksg_fast_count": (C.c_int, [_c_prep, _c_i64, _c_i64, C.POINTER(C.c_uint32), C.c_int, _c_vp, C.c_int,
                                 C.c_int32, C.c_int, _c_vp, _c_vp, _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_axis_count": (C.c_int, [_c_prep, C.c_int, _c_i64, _c_i64, _c_vp, C.c_int, C.c_int32, _c_vp, _c_vp]),
    "ksg_axis_count_multi": (C.c_int, [_c_prep, C.POINTER(C.c_int32), C.c_int, _c_i64, _c_i64, _c_vp, C.c_int, C.c_int32,
                                       _c_vp, _c_vp]),
    "ksg_axis_count_buckets": (C.c_int, [_c_prep, C.POINTER(C.c_int32), C.c_int, _c_i64, _c_i64, _c_vp, C.c_int, C.c_int32,
                                         _c_vp, _c_vp, _c_vp]),
    "ksg_prepare_axes": (C.c_int, [_c_prep, _c_vp]),
    "ksg_finish_workspace": (C.c_size_t, [_c_i64]),
    "ksg_psi_finish": (C.c_int, [_c_vp, _c_i64, C.c_int, C.POINTER(PsiStep), C.c_int, C.c_double, C.c_double, C.c_int,
                                 C.c_double, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_scatter_sum_f64": (C.c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_psi_epilogue_push": (C.c_int, [_c_vp, _c_i64, C.c_int, C.POINTER(PsiStep), C.c_int, C.c_double, C.c_double, C.c_int,
                                        C.c_double, _c_vp, _c_i64, C.POINTER(C.c_void_p), _c_i64, _c_vp,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_ulonglong,
                                        _c_vp, _c_vp]),
    "ksg_p2p_push": (C.c_int, [_c_vp, C.c_size_t, C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(C.c_void_p), C.c_int,
                               C.c_int, C.c_ulonglong, _c_vp, _c_vp]),
    "ksg_p2p_wait": (C.c_int, [_c_vp, C.c_int, C.c_ulonglong, _c_vp, _c_vp, _c_vp]),
    "ksg_scatter_f64": (C.c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp]),
    "ksg_scatter_i32": (C.c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp]),
    "ksg_gather_f64": (C.c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp]),
    "ksg_sum_i32": (C.c_int, [_c_vp, _c_i64, _c_vp, _c_vp]),
    "ksg_knn_radius_p": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, _c_vp, _c_i64, _c_i64, _c_i64, C.c_int, C.c_double,
                                   _c_vp, _c_vp, _c_vp, C.c_size_t, _c_vp]),
    "ksg_count_subspaces_p": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, _c_vp, _c_i64, _c_i64, _c_i64,
                                        C.POINTER(C.c_uint32), C.c_int, _c_vp, _c_i64, C.c_int, C.c_int32, C.c_double,
                                        _c_vp, _c_vp]),
    "ksg_subspace_radius_from_neighbours_p": (C.c_int, [_c_vp, _c_i64, _c_i64, C.c_int, _c_i64, _c_i64, _c_vp,
                                                        C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int, C.c_double,
                                                        _c_vp, _c_vp]),
    "ksg_paircount_fixed_radius_p": (C.c_int, [_c_vp, _c_vp, _c_i64, C.c_int, C.c_double, C.c_double, _c_i64, _c_i64,
                                               _c_vp, _c_vp]),
    "ksg_sort_workspace": (C.c_size_t, [_c_i64]),
    "ksg_sort_f64": (C.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, C.c_int, _c_vp, C.c_size_t, _c_vp]),
}

_lib = None


class KsgError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building in-tree first if sources are newer) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this host: use the shipped binary if any
            if not os.path.exists(path):
                raise ImportError(
                    "libksg_b200.so is missing and could not be built; the KSG path has no "
                    "CPU fallback (%s)" % exc
                ) from exc
            import warnings

            warnings.warn("libksg_b200.so is older than its sources and could not be rebuilt (%s): loading the "
                          "existing binary; header / ABI changes since then are NOT in it" % exc, RuntimeWarning)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ksg_last_error().decode("utf-8", "replace")
        raise KsgError(f"{what} failed with code {rc}: {msg}")
