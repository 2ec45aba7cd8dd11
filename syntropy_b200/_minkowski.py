"""Estimator drivers for Minkowski p-norms other than the maximum norm.

The reference forwards ``p`` / ``norm`` to scipy's cKDTree (knn/utils.py:33,81; knn/shannon.py:82,
186-192,247-262,332-338,401-402; knn/temporal.py:71-72,142-143); the device side is the exact float64
dense kernel family of ``csrc/minkowski_f64.cuh`` (p = 1, 2: bit-exact; other p: device ``pow``).
Everything runs in the caller's sample order; query rows shard over the ranks like the exact engine.
``p = inf`` never comes here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from . import _engine as E
from . import dist as _dist


def check_p(p) -> float:
    """Validate a finite p (cKDTree raises the same ValueError for p < 1)."""
    p = float(p)
    if not p >= 1.0:
        raise ValueError("Only p-norms with 1<=p<=infinity permitted")
    return p


def knn_radius_p(points: torch.Tensor, kprime: int, q_begin: int, q_end: int, p: float,
                 queries: torch.Tensor | None = None, want_indices: bool = False):
    lib = _cabi.load()
    if kprime > E.MAX_KPRIME:
        raise NotImplementedError(f"k + 1 = {kprime} exceeds the neighbour-list limit of the CUDA kernels ({E.MAX_KPRIME})")
    dim, n = points.shape
    nq = q_end - q_begin
    eps = torch.empty(nq, dtype=torch.float64, device=points.device)
    idx = torch.empty((nq, kprime), dtype=torch.int64, device=points.device) if want_indices else None
    ws = torch.empty(max(int(lib.ksg_knn_radius_workspace(n, max(nq, 1), dim, kprime)), 1), dtype=torch.uint8,
                     device=points.device)
    with E.timer.span("knn_radius_p"):
        rc = lib.ksg_knn_radius_p(E._ptr(points), points.stride(0), n, dim, E._ptr(queries),
                                  0 if queries is None else queries.stride(0), q_begin, q_end, kprime, float(p),
                                  E._ptr(eps), E._ptr(idx), E._ptr(ws), ws.numel(), E._stream())
    _cabi.check(rc, "ksg_knn_radius_p")
    return (eps, idx) if want_indices else eps


def count_subspaces_p(points: torch.Tensor, masks, eps: torch.Tensor, q_begin: int, q_end: int, p: float,
                      strict: bool = True, bias: int = -1, per_subspace_eps: bool = False) -> torch.Tensor:
    lib = _cabi.load()
    dim, n = points.shape
    nq = q_end - q_begin
    counts = torch.empty((len(masks), nq), dtype=torch.int32, device=points.device)
    with E.timer.span("count_subspaces_p"):
        rc = lib.ksg_count_subspaces_p(E._ptr(points), points.stride(0), n, dim, None, 0, q_begin, q_end,
                                       E._masks(masks), len(masks), E._ptr(eps), nq if per_subspace_eps else 0,
                                       1 if strict else 0, bias, float(p), E._ptr(counts), E._stream())
    _cabi.check(rc, "ksg_count_subspaces_p")
    return counts


def _chebyshev_radius(pts: torch.Tensor, k: int, q0: int, q1: int) -> torch.Tensor:
    """Maximum-norm joint radii of samples [q0, q1) in the caller's order (``mutual_information_1``
    issues its joint query without ``p``, knn/shannon.py:186): the fast engine when it applies."""
    dim, n = pts.shape
    if E._use_fast(dim, k + 1) and _dist.world()[1] == 1:
        prep = E.PreparedSet(pts, E.order_for(dim))
        eps_sorted = E.fast_knn(prep, k + 1, 0, n, E.config.engine == "windowed")
        return E.scatter_to_original(eps_sorted, prep)[q0:q1].contiguous()
    return E.knn_radius(pts, k + 1, q0, q1)


def entropy(rows, k: int, p: float):
    """knn/shannon.py:81-87 with the (k+1)-th neighbour distance under the p-norm (:82)."""
    p = check_p(p)
    pts = E.to_device_points(rows)
    dim, n = pts.shape
    status = E.Status(pts.device)
    E.check_finite(pts, out=status.finite)
    q0, q1 = _dist.my_slice(n)
    eps = knn_radius_p(pts, k + 1, q0, q1, p)
    ptw = torch.empty(q1 - q0, dtype=torch.float64, device=pts.device)
    rc = _cabi.load().ksg_entropy_epilogue(E._ptr(eps), q1 - q0, dim, E.psi_host(k), E.psi_host(n), E._ptr(ptw), E._stream())
    _cabi.check(rc, "ksg_entropy_epilogue")
    return E._finish(E.Pointwise(ptw), n, status)


def ksg1(rows, k: int, masks, prog_builder, p: float, joint_chebyshev: bool):
    """Joint radius -> strict counts under the p-norm in every subspace -> psi program.
    ``joint_chebyshev``: the joint query ignores ``p`` (mutual_information_1, knn/shannon.py:186 vs
    :192); the conditional MI forwards it to both (:332,338)."""
    p = check_p(p)
    pts = E.to_device_points(rows)
    dim, n = pts.shape
    status = E.Status(pts.device)
    E.check_finite(pts, out=status.finite)
    q0, q1 = _dist.my_slice(n)
    prog = prog_builder(E.psi_host(k), E.psi_host(n))
    eps = _chebyshev_radius(pts, k, q0, q1) if joint_chebyshev else knn_radius_p(pts, k + 1, q0, q1, p)
    counts = count_subspaces_p(pts, masks, eps, q0, q1, p, strict=True, bias=-1)
    return E._finish(E.Pointwise(E.psi_epilogue(counts, prog, n)), n, status)


def ksg2(rows, k: int, masks, prog_builder, p: float):
    """KSG algorithm 2 under the p-norm (knn/shannon.py:247-263): neighbours of the joint query,
    per-marginal radius = largest p-norm distance to them, closed-ball counts, psi(count)."""
    p = check_p(p)
    lib = _cabi.load()
    pts = E.to_device_points(rows)
    dim, n = pts.shape
    status = E.Status(pts.device)
    E.check_finite(pts, out=status.finite)
    q0, q1 = _dist.my_slice(n)
    nq = q1 - q0
    prog = prog_builder(E.psi_host(k), E.psi_host(n))
    _, idx = knn_radius_p(pts, k + 1, q0, q1, p, want_indices=True)
    eps_s = torch.empty((len(masks), nq), dtype=torch.float64, device=pts.device)
    rc = lib.ksg_subspace_radius_from_neighbours_p(E._ptr(pts), pts.stride(0), n, dim, q0, q1, E._ptr(idx), k + 1, 1,
                                                   E._masks(masks), len(masks), float(p), E._ptr(eps_s), E._stream())
    _cabi.check(rc, "ksg_subspace_radius_from_neighbours_p")
    counts = count_subspaces_p(pts, masks, eps_s, q0, q1, p, strict=False, bias=-1, per_subspace_eps=True)
    return E._finish(E.Pointwise(E.psi_epilogue(counts, prog, n)), n, status)


def kl_divergence(post_rows, prior_rows, k: int, p: float):
    """knn/shannon.py:401-407 with both neighbour searches under the p-norm."""
    p = check_p(p)
    lib = _cabi.load()
    P = E.to_device_points(post_rows)
    Q = E.to_device_points(prior_rows)
    dim, n = P.shape
    m = Q.shape[1]
    status = E.Status(P.device)
    E.check_finite(P, out=status.finite)
    E.check_finite(Q, out=status.finite)
    q0, q1 = _dist.my_slice(n)
    with np.errstate(all="ignore"):
        log_ratio = float(np.log(m / (n - 1)))
    rho = knn_radius_p(P, k + 1, q0, q1, p)
    nu = knn_radius_p(Q, k, q0, q1, p, queries=P)
    ptw = torch.empty(q1 - q0, dtype=torch.float64, device=P.device)
    rc = lib.ksg_kl_epilogue(E._ptr(nu), E._ptr(rho), q1 - q0, dim, log_ratio, E._ptr(ptw), E._stream())
    _cabi.check(rc, "ksg_kl_epilogue")
    return E._finish(E.Pointwise(ptw), n, status)


def pair_counts(x: np.ndarray, y: np.ndarray, m: int, r: float, p: float):
    """(small, big) ordered pair counts of the m / (m+1) templates within p-norm distance r
    (cKDTree.count_neighbors(.., r, p=norm), knn/temporal.py:71-72,142-143)."""
    p = check_p(p)
    dev = E.require_cuda()
    xs = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    ys = xs if y is x else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
    nt = xs.numel() - m
    if nt <= 0:
        raise ValueError("series shorter than the template length")
    flag = E.check_finite(xs) | E.check_finite(ys)
    i0, i1 = _dist.my_slice(nt)
    out = torch.zeros(2, dtype=torch.int64, device=dev)
    rc = _cabi.load().ksg_paircount_fixed_radius_p(E._ptr(xs), E._ptr(ys), nt, m, float(r), float(p), i0, i1,
                                                   E._ptr(out), E._stream())
    _cabi.check(rc, "ksg_paircount_fixed_radius_p")
    out = _dist.sum_counts(out)
    host = torch.cat([out, flag.to(torch.int64)]).cpu().numpy()
    if host[2] != 0:
        raise ValueError("data must be finite, check for nan or inf values")
    return int(host[0]), int(host[1])
