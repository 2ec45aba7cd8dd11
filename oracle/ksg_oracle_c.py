"""ctypes wrapper of oracle/ksg_oracle.c -- TEST INFRASTRUCTURE ONLY (see ksg_oracle.py).
Used where the numpy oracle would be slow (query subsets against multi-million point sets)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "_build", "libksg_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "ksg_oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(_LIB)
    return _lib


def set_threads(n: int = 0) -> int:
    """Use ``n`` OpenMP threads from now on (0: keep the default); returns the count in effect."""
    return int(load().oracle_set_threads(C.c_int(int(n))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def knn_radius(points, kprime, queries=None):
    """kprime-th smallest Chebyshev distance of each query to the points (no implicit +1)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    qs = pts if queries is None else np.ascontiguousarray(queries, dtype=np.float64)
    dim, n = pts.shape
    nq = qs.shape[1]
    out = np.empty(nq)
    load().oracle_knn_radius(_p(pts), C.c_int64(n), C.c_int64(n), C.c_int(dim), _p(qs), C.c_int64(nq),
                             C.c_int64(nq), C.c_int(kprime), _p(out))
    return out


def count_within(points, eps, strict=True, queries=None, bias=-1):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    qs = pts if queries is None else np.ascontiguousarray(queries, dtype=np.float64)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    dim, n = pts.shape
    nq = qs.shape[1]
    out = np.empty(nq, dtype=np.int64)
    load().oracle_count_within(_p(pts), C.c_int64(n), C.c_int64(n), C.c_int(dim), _p(qs), C.c_int64(nq),
                               C.c_int64(nq), _p(eps), C.c_int(1 if strict else 0), C.c_int64(bias), _p(out))
    return out


def pair_counts(x, y, m, r):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    nt = x.size - m
    out = np.zeros(2, dtype=np.int64)
    load().oracle_pair_count(_p(x), _p(y), C.c_int64(nt), C.c_int(m), C.c_double(r), _p(out))
    return int(out[0]), int(out[1])


def knn_radius_p(points, kprime, p, queries=None, return_indices=False):
    """kprime-th smallest Minkowski-p distance (1 <= p < inf) of each query to the points."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    qs = pts if queries is None else np.ascontiguousarray(queries, dtype=np.float64)
    dim, n = pts.shape
    nq = qs.shape[1]
    out = np.empty(nq)
    idx = np.empty((nq, kprime), dtype=np.int64) if return_indices else None
    load().oracle_knn_radius_p(_p(pts), C.c_int64(n), C.c_int64(n), C.c_int(dim), _p(qs), C.c_int64(nq),
                               C.c_int64(nq), C.c_int(kprime), C.c_double(float(p)), _p(out),
                               _p(idx) if return_indices else None)
    return (out, idx) if return_indices else out


def count_within_p(points, eps, p, strict=True, queries=None, bias=-1):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    qs = pts if queries is None else np.ascontiguousarray(queries, dtype=np.float64)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    dim, n = pts.shape
    nq = qs.shape[1]
    out = np.empty(nq, dtype=np.int64)
    load().oracle_count_within_p(_p(pts), C.c_int64(n), C.c_int64(n), C.c_int(dim), _p(qs), C.c_int64(nq),
                                 C.c_int64(nq), _p(eps), C.c_int(1 if strict else 0), C.c_int64(bias),
                                 C.c_double(float(p)), _p(out))
    return out


def pair_count_p(u, v, r, p):
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    fn = load().oracle_pair_count_p
    fn.restype = C.c_int64
    return int(fn(_p(u), C.c_int64(u.shape[1]), C.c_int64(u.shape[1]), _p(v), C.c_int64(v.shape[1]),
                  C.c_int64(v.shape[1]), C.c_int(u.shape[0]), C.c_double(float(r)), C.c_double(float(p))))
