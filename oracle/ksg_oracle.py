"""CPU oracle for the syntropy.knn KSG hot path -- TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs may
import it.  ``syntropy_b200`` never imports anything under ``oracle/``.

What it restates
----------------
The reference (``/root/reference/syntropy/knn``) is pure Python; the arithmetic of
the path lives in a third-party dependency that is NOT vendored in the reference
tree: ``scipy.spatial.cKDTree`` (query / query_ball_point / count_neighbors) and
``scipy.special.digamma`` (pin in the reference: ``scipy>=1.7.0``,
``pyproject.toml:36``; installed in this image: scipy 1.18.1).  A kd-tree is only
an accelerator for an exact query, so the published algorithm is restated here
tree-free, as brute force over all pairs in float64:

* radius   ``eps_i   = (k+1)-th smallest over ALL j (self included) of
                        max_d |fl64(x_id - x_jd)|``            (knn/utils.py:29-33)
* count    ``c_i     = #{j : max_{d in s} |fl64(x_id - x_jd)| <  eps_i} - 1``   strict
           ``c_i     = #{j : ...                              <= eps_i} - 1``   closed
                                                               (knn/utils.py:75-84)
* pairs    ``#{(i,j) : max_d |u_id - v_jd| <= r}``             (knn/temporal.py:71-72,142-143)
* psi      cephes ``psi`` on positive integers: harmonic sum - EULER for n <= 10,
           ``log(x) - 0.5/x - z*polevl(z, A)`` with z = 1/x^2 above
           (scipy.special.digamma at the reference's call sites)

Parity pinning (see tests/test_oracle.py): the estimators below reproduce all nine
known answers of the reference's ``tests/test_knn.py:29-75``, the hand-computed
radii / strict-vs-closed counts of ``tests/test_knn_utils.py:10-92``, the ten
Antropy constants of ``tests/test_knn_temporal.py:55-76`` and the committed
fixtures under ``tests/golden/`` that were produced by running the unmodified
reference in the build container (``tests/golden/make_golden.py``).

Every estimator follows the float64 operation order of the reference line by line
so that pointwise arrays agree to the last bit when the psi values do.
"""
from __future__ import annotations

import math
import numpy as np

EULER = 0.57721566490153286061

# cephes psi asymptotic-series coefficients (Bernoulli numbers B_2k / 2k), as
# published in cephes/psi.c and xsf/digamma.h
_PSI_A = (
    8.33333333333333333333e-2,
    -2.10927960927960927961e-2,
    7.57575757575757575758e-3,
    -4.16666666666666666667e-3,
    3.96825396825396825397e-3,
    -8.33333333333333333333e-3,
    8.33333333333333333333e-2,
)


def psi_int(n):
    """digamma on integer arguments, float64 -- restates cephes ``psi`` for the
    only arguments the path ever uses (counts+1, k, N).

    n <= 0 -> -inf (the reference reaches psi(0) when eps == 0, Appendix A).
    """
    n = np.asarray(n, dtype=np.int64)
    out = np.empty(n.shape, dtype=np.float64)
    flat_n = n.ravel()
    flat_o = out.ravel()
    small = flat_n <= 10
    # harmonic numbers for 1..10, summed in cephes order: y = sum_{i=1}^{n-1} 1/i ; y -= EULER
    table = np.empty(11, dtype=np.float64)
    table[0] = -np.inf
    for m in range(1, 11):
        y = 0.0
        for i in range(1, m):
            y += 1.0 / i
        table[m] = y - EULER
    idx = np.clip(flat_n[small], 0, 10)
    flat_o[small] = table[idx]
    big = ~small
    if big.any():
        s = flat_n[big].astype(np.float64)
        z = 1.0 / (s * s)
        poly = np.full_like(z, _PSI_A[0])
        for c in _PSI_A[1:]:
            poly = poly * z + c
        y = z * poly
        flat_o[big] = np.log(s) - 0.5 / s - y
    return out


def _as_f64(points):
    a = np.ascontiguousarray(np.asarray(points), dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("points must be (n_dims, n_samples)")
    if not np.isfinite(a).all():
        raise ValueError("data must be finite")  # cKDTree raises ValueError too
    return a


def _cheb_block(q, pts):
    """(nq, n) matrix of max_d |q_d - p_d| in float64; q (D,nq), pts (D,n)."""
    dist = np.abs(q[0][:, None] - pts[0][None, :])
    for d in range(1, q.shape[0]):
        np.maximum(dist, np.abs(q[d][:, None] - pts[d][None, :]), out=dist)
    return dist


# ---- Minkowski p-norms other than the maximum norm.  cKDTree works in
# "p-th power space" (scipy/spatial/ckdtree/src/distance_base.h, distance.h -- not vendored in the
# reference; the published algorithm, pinned below against the installed scipy):
#   p = 1    s = sum_d |x_d - y_d|                          sequential over d; distance = s
#   p = 2    s = sum_d (x_d - y_d)^2                        four partial sums over d = 0,4,8.. / 1,5,.. /
#                                                           2,6,.. / 3,7,.., added as ((a0+a1)+a2)+a3, then
#                                                           the d % 4 remainder sequentially; distance = sqrt(s)
#   other p  s = sum_d pow(|x_d - y_d|, p)                  sequential; distance = pow(s, 1/p)
# and a ball query of radius r keeps s <= r (p = 1), s <= r*r (p = 2), s <= pow(r, p) (other p).
def _is_inf(p):
    return p == np.inf


def _pow_dist_block(q, pts, p):
    """(nq, n) matrix of the p-th-power-space distances, operation order as above."""
    D = q.shape[0]
    diff = [q[d][:, None] - pts[d][None, :] for d in range(D)]
    if p == 1:
        s = np.abs(diff[0])
        for d in range(1, D):
            s = s + np.abs(diff[d])
        return s
    if p == 2:
        full = D - D % 4
        if full:
            acc = [np.zeros_like(diff[0]) for _ in range(4)]
            for d in range(0, full, 4):
                for u in range(4):
                    acc[u] = acc[u] + diff[d + u] * diff[d + u]
            s = ((acc[0] + acc[1]) + acc[2]) + acc[3]
        else:
            s = np.zeros_like(diff[0])
        for d in range(full, D):
            s = s + diff[d] * diff[d]
        return s
    s = np.power(np.abs(diff[0]), p)
    for d in range(1, D):
        s = s + np.power(np.abs(diff[d]), p)
    return s


def _root_p(s, p):
    if p == 1:
        return s
    if p == 2:
        return np.sqrt(s)
    with np.errstate(all="ignore"):
        return np.power(s, 1.0 / p)


def _bound_p(r, p):
    r = np.asarray(r, dtype=np.float64)
    if p == 1:
        return r
    if p == 2:
        return r * r
    with np.errstate(all="ignore"):
        return np.power(r, p)


def _check_p(p):
    if not (p >= 1):
        raise ValueError("Only p-norms with 1<=p<=infinity permitted")  # what cKDTree raises


def knn_radius_p(points, k, p, queries=None, block=256, return_indices=False):
    """``knn_radius`` under the Minkowski p-norm (1 <= p < inf), cKDTree's arithmetic."""
    _check_p(p)
    pts = _as_f64(points)
    n = pts.shape[1]
    if queries is None:
        qs, kp = pts, k + 1
    else:
        qs, kp = _as_f64(queries), k
    if p not in (1, 2):
        # numpy's vectorised power is not the libm pow() scipy's C++ calls: the C restatement is
        from . import ksg_oracle_c as oc
        return oc.knn_radius_p(pts, kp, p, queries=qs, return_indices=return_indices)
    nq = qs.shape[1]
    eps = np.full(nq, np.inf)
    idx = np.full((nq, kp), n, dtype=np.int64)
    for b in range(0, nq, block):
        s = _pow_dist_block(qs[:, b:b + block], pts, p)
        if kp <= n:
            order = np.argsort(s, axis=1, kind="stable")[:, :kp]
            idx[b:b + block] = order
            eps[b:b + block] = _root_p(np.take_along_axis(s, order[:, -1:], axis=1)[:, 0], p)
    return (eps, idx) if return_indices else eps


def count_within_p(points, eps, p, strict=True, queries=None, block=256):
    """``count_within`` under the Minkowski p-norm: ``#{j : s_ij <= bound(r_i)} - 1`` with
    ``r = nextafter(eps, -inf)`` when strict (knn/utils.py:75-84 with ``p`` forwarded)."""
    _check_p(p)
    pts = _as_f64(points)
    qs = pts if queries is None else _as_f64(queries)
    eps = np.asarray(eps, dtype=np.float64)
    if p not in (1, 2):
        from . import ksg_oracle_c as oc
        return oc.count_within_p(pts, eps, p, strict=strict, queries=qs)
    r = np.nextafter(eps, -np.inf) if strict else eps
    ub = _bound_p(r, p)
    nq = qs.shape[1]
    out = np.empty(nq, dtype=np.int64)
    for b in range(0, nq, block):
        s = _pow_dist_block(qs[:, b:b + block], pts, p)
        out[b:b + block] = (s <= ub[b:b + block][:, None]).sum(axis=1) - 1
    return out


def pair_count_p(u, v, r, p, block=512):
    """``pair_count`` under the Minkowski p-norm (cKDTree.count_neighbors(.., r, p))."""
    _check_p(p)
    u, v = _as_f64(u), _as_f64(v)
    if p not in (1, 2):
        from . import ksg_oracle_c as oc
        return oc.pair_count_p(u, v, r, p)
    ub = float(_bound_p(r, p))
    total = 0
    for b in range(0, u.shape[1], block):
        total += int((_pow_dist_block(u[:, b:b + block], v, p) <= ub).sum())
    return total


def knn_radius(points, k, queries=None, block=256, return_indices=False):
    """k'-th smallest Chebyshev distance from each query to the point set.

    Within-set use (queries=None): k' = k+1 with self included at distance 0,
    exactly what ``tree.query(data.T, k=k+1)`` returns in ``distances[:, -1]``
    (knn/utils.py:33).  If k' exceeds the set size the radius is +inf (cKDTree
    pads missing neighbours with inf).
    Cross-set use (queries given): k' = k, no self (knn/shannon.py:402).
    """
    pts = _as_f64(points)
    n = pts.shape[1]
    if queries is None:
        qs, kp = pts, k + 1
    else:
        qs, kp = _as_f64(queries), k
    nq = qs.shape[1]
    eps = np.full(nq, np.inf)
    idx = np.full((nq, kp), n, dtype=np.int64)
    for b in range(0, nq, block):
        dist = _cheb_block(qs[:, b:b + block], pts)
        if kp <= n:
            if return_indices:
                order = np.argsort(dist, axis=1, kind="stable")[:, :kp]
                idx[b:b + block] = order
                eps[b:b + block] = np.take_along_axis(dist, order[:, -1:], axis=1)[:, 0]
            else:
                eps[b:b + block] = np.partition(dist, kp - 1, axis=1)[:, kp - 1]
    if return_indices:
        return eps, idx
    return eps


def count_within(points, eps, strict=True, queries=None, block=256):
    """Neighbour count inside a per-query radius, self removed by ``- 1``.

    Restates knn/utils.py:75-84: the reference nudges the radius down with
    ``nextafter`` and asks for the closed ball, which is the open ball
    ``dist < eps``; the query point itself is inside the ball (and removed by the
    ``- 1``) -- also when eps == 0, where the strict ball is empty and the result
    is -1 (Appendix A).
    """
    pts = _as_f64(points)
    qs = pts if queries is None else _as_f64(queries)
    eps = np.asarray(eps, dtype=np.float64)
    nq = qs.shape[1]
    out = np.empty(nq, dtype=np.int64)
    for b in range(0, nq, block):
        dist = _cheb_block(qs[:, b:b + block], pts)
        e = eps[b:b + block][:, None]
        hit = (dist < e) if strict else (dist <= e)
        out[b:b + block] = hit.sum(axis=1) - 1
    return out


def pair_count(u, v, r, block=512):
    """#ordered pairs (i,j) with max_d |u_id - v_jd| <= r  (closed ball).

    Restates ``tree_u.count_neighbors(tree_v, r, p=inf)`` (knn/temporal.py:71-72,
    142-143); self pairs are included, as there.
    """
    u = _as_f64(u)
    v = _as_f64(v)
    total = 0
    for b in range(0, u.shape[1], block):
        total += int((_cheb_block(u[:, b:b + block], v) <= r).sum())
    return total


def check_idxs(idxs, data):
    # syntropy/utils/utils.py:9-35
    return tuple(range(np.asarray(data).shape[0])) if idxs is None else tuple(idxs)


def _rows(data, idxs):
    return np.asarray(data)[tuple(idxs), :]


# --------------------------------------------------------------------------- estimators

def differential_entropy(data, k, idxs=None, p=np.inf, noise_level=0.0, seed=0):
    """knn/shannon.py:11-87 (Kozachenko-Leonenko)."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    d = len(idxs_)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    sub = _rows(data, idxs_)
    if noise_level > 0:
        # note: the reference adds float noise IN PLACE to the fancy-index copy
        # (knn/shannon.py:80); an integer copy would raise there, as it does here.
        rng = np.random.default_rng(seed)
        sub += rng.standard_normal(sub.shape) * noise_level
    eps = knn_radius(sub, k) if _is_inf(p) else knn_radius_p(sub, k, p)   # knn/shannon.py:82
    ptw = np.zeros((1, N))
    with np.errstate(divide="ignore"):
        ptw[0, :] += -psi_k + psi_N + d * np.log(2 * eps)
    return ptw, ptw.mean()


def mutual_information(idxs_x, idxs_y, k, data, algorithm=1, p=np.inf):
    """knn/shannon.py:90-134 dispatcher."""
    assert algorithm in {1, 2}, "Algorithm must be 1 or 2."
    if algorithm == 1:
        return mutual_information_1(idxs_x, idxs_y, k, data, p=p)
    return mutual_information_2(idxs_x, idxs_y, k, data, p=p)


def mutual_information_1(idxs_x, idxs_y, k, data, p=np.inf):
    """knn/shannon.py:137-196 (KSG-1).  ``p`` reaches the marginal counts only: the joint query is
    issued without it (:186, maximum norm), the counts forward it (:192)."""
    data = np.asarray(data)
    idxs_xy = tuple(idxs_x) + tuple(idxs_y)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    eps = knn_radius(_rows(data, idxs_xy), k)
    ptw = np.full((1, N), psi_k + psi_N)
    for idxs in (idxs_x, idxs_y):
        counts = count_within(_rows(data, idxs), eps) if _is_inf(p) else count_within_p(_rows(data, idxs), eps, p)
        ptw[0, :] -= psi_int(counts + 1)
    return ptw, ptw.mean()


def _ksg2_subspace_radius(sub, neighbours, p=np.inf):
    """eps_s = max_j<=k ||x_i - x_nbr_j||_p in the subspace, ``np.linalg.norm(.., ord=p, axis=1)``
    (knn/shannon.py:253-259): |.|.max() for inf, sum |.| for 1, sqrt(sum .^2) for 2 (left-to-right
    over the < 8 coordinates), (sum |.|^p)^(1/p) otherwise."""
    N = sub.shape[1]
    eps = np.full(N, -np.inf)
    for j in range(neighbours.shape[1]):
        diff = (sub - sub[:, neighbours[:, j]]).T        # (N, d), as the reference forms it
        eps = np.maximum(np.linalg.norm(diff, ord=p, axis=1), eps)
    return eps


def mutual_information_2(idxs_x, idxs_y, k, data, p=np.inf):
    """knn/shannon.py:199-265 (KSG-2).  Neighbour identity among exact distance
    ties is cKDTree-internal; this restatement breaks ties by index."""
    data = np.asarray(data)
    idxs_xy = tuple(idxs_x) + tuple(idxs_y)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    if _is_inf(p):
        _, indices = knn_radius(_rows(data, idxs_xy), k, return_indices=True)
    else:
        _, indices = knn_radius_p(_rows(data, idxs_xy), k, p, return_indices=True)
    neighbours = indices[:, 1:]
    ptw = np.full((1, N), psi_k - (1 / k) + psi_N)
    for idxs in (idxs_x, idxs_y):
        sub = _as_f64(_rows(data, idxs))
        eps = _ksg2_subspace_radius(sub, neighbours, p)
        counts = count_within(sub, eps, strict=False) if _is_inf(p) else count_within_p(sub, eps, p, strict=False)
        ptw[0, :] -= psi_int(counts)
    return ptw, ptw.mean()


def conditional_mutual_information(idxs_x, idxs_y, idxs_z, k, data, p=np.inf):
    """knn/shannon.py:268-347 (Frenzel-Pompe)."""
    data = np.asarray(data)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    idxs_x, idxs_y, idxs_z = tuple(idxs_x), tuple(idxs_y), tuple(idxs_z)
    ptw = np.full((1, N), psi_k)
    joint = _rows(data, idxs_x + idxs_y + idxs_z)
    eps = knn_radius(joint, k) if _is_inf(p) else knn_radius_p(joint, k, p)   # :332, p forwarded
    for c, idxs in enumerate((idxs_z, idxs_x + idxs_z, idxs_y + idxs_z)):
        counts = count_within(_rows(data, idxs), eps) if _is_inf(p) else count_within_p(_rows(data, idxs), eps, p)
        if c == 0:
            ptw[0, :] += psi_int(counts + 1)
        else:
            ptw[0, :] -= psi_int(counts + 1)
    return ptw, ptw.mean()


def kullback_leibler_divergence(posterior_data, prior_data, k, p=np.inf):
    """knn/shannon.py:350-410 (Wang-Kulkarni-Verdu / Perez-Cruz)."""
    P = np.asarray(posterior_data)
    Q = np.asarray(prior_data)
    assert P.shape[0] == Q.shape[0], "The data must have the same number of dimensions."
    d, n = P.shape
    m = Q.shape[1]
    if _is_inf(p):
        rho = knn_radius(P, k)
        nu = knn_radius(Q, k, queries=P)
    else:
        rho = knn_radius_p(P, k, p)                # :401
        nu = knn_radius_p(Q, k, p, queries=P)      # :402
    ptw = d * np.log(nu / rho) + np.log(m / (n - 1))
    return ptw, np.mean(ptw)


def total_correlation(data, k, idxs=None, algorithm=1):
    """knn/multivariate_mi.py:11-51 dispatcher."""
    assert algorithm in {1, 2}, "Algorithm must be 1 or 2."
    if algorithm == 1:
        return total_correlation_1(data, k, idxs)
    return total_correlation_2(data, k, idxs)


def total_correlation_1(data, k, idxs=None):
    """knn/multivariate_mi.py:54-115."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    ptw = np.full((1, N), psi_k + (m - 1) * psi_N)
    eps = knn_radius(_rows(data, idxs_), k)
    for idx in idxs_:
        counts = count_within(_rows(data, (idx,)), eps)
        ptw[0, :] -= psi_int(counts + 1)
    return ptw, ptw.mean()


def total_correlation_2(data, k, idxs=None):
    """knn/multivariate_mi.py:118-193."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    ptw = np.full((1, N), psi_k - ((m - 1) / k) + ((m - 1) * psi_N))
    _, indices = knn_radius(_rows(data, idxs_), k, return_indices=True)
    neighbours = indices[:, 1:]
    for idx in idxs_:
        sub = _as_f64(_rows(data, (idx,)))
        eps = _ksg2_subspace_radius(sub, neighbours)
        counts = count_within(sub, eps, strict=False)
        ptw[0, :] -= psi_int(counts)
    return ptw, ptw.mean()


def dual_total_correlation(data, k, idxs=None):
    """knn/multivariate_mi.py:196-260."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    N = data.shape[1]
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    ptw = np.full((1, N), psi_k - psi_N)
    eps = knn_radius(_rows(data, idxs_), k)
    for i in range(m):
        rest = [idxs_[j] for j in range(m) if j != i]
        counts = count_within(_rows(data, rest), eps)
        ptw[0, :] -= (psi_int(counts + 1) - psi_N) / (m - 1)
    ptw *= m - 1
    return ptw, ptw.mean()


def s_information(data, k, idxs=None):
    """knn/multivariate_mi.py:263-335."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    N = data.shape[1]
    m = len(idxs_)
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    eps = knn_radius(_rows(data, idxs_), k)
    ptw = np.full((1, N), psi_k - psi_N)
    for d in range(m):
        small = data[idxs_[d]:idxs_[d] + 1, :]
        big = _rows(data, [idxs_[j] for j in range(m) if j != d])
        counts_small = count_within(small, eps)
        counts_big = count_within(big, eps)
        ptw[0, :] -= (psi_int(counts_big + 1) - psi_N) / m
        ptw[0, :] -= (psi_int(counts_small + 1) - psi_N) / m
    ptw *= m
    return ptw, ptw.mean()


def o_information(data, k, idxs=None):
    """knn/multivariate_mi.py:338-415."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    N = data.shape[1]
    m = len(idxs_)
    if m == 2:
        return np.zeros(N), 0.0
    psi_k = float(psi_int(k))
    psi_N = float(psi_int(N))
    eps = knn_radius(_rows(data, idxs_), k)
    ptw = np.full((1, N), psi_k - psi_N)
    for d in range(m):
        small = data[idxs_[d]:idxs_[d] + 1, :]
        big = _rows(data, [idxs_[j] for j in range(m) if j != d])
        counts_small = count_within(small, eps)
        counts_big = count_within(big, eps)
        ptw[0, :] -= (psi_int(counts_big + 1) - psi_N) / (m - 2)
        ptw[0, :] += (psi_int(counts_small + 1) - psi_N) / (m - 2)
    ptw *= 2 - m
    return ptw, ptw.mean()


def _embed(series, width, count):
    """(width, count) delay embedding: row w is series[w : w + count]
    (``sliding_window_view(series, width)[:count]`` transposed)."""
    return np.stack([series[w:w + count] for w in range(width)])


def sample_entropy(idx, data, m=2, r=0.2, normalize_r=True, norm=np.inf):
    """knn/temporal.py:10-77."""
    series = np.asarray(data)[idx[0], :].astype(np.float64)
    N = series.shape[0] - m
    small = _embed(series, m, N)
    big = _embed(series, m + 1, N)
    if normalize_r:
        r = r * np.std(series)
    pc = pair_count if _is_inf(norm) else (lambda u, v, rr: pair_count_p(u, v, rr, norm))   # :71-72
    counts_small = pc(small, small, r) - N
    counts_big = pc(big, big, r) - N
    if counts_big == 0:
        return np.inf
    return -np.log(counts_big / counts_small)


def zscore(x):
    """scipy.stats.zscore defaults (ddof=0): (x - mean) / std."""
    x = np.asarray(x, dtype=np.float64)
    return (x - x.mean()) / x.std()


def cross_sample_entropy(idx_x, idx_y, data, m=2, r=0.2, norm=np.inf):
    """knn/temporal.py:79-148."""
    data = np.asarray(data)
    N = data.shape[1] - m
    sx = zscore(data[idx_x[0], :])
    sy = zscore(data[idx_y[0], :])
    pc = pair_count if _is_inf(norm) else (lambda u, v, rr: pair_count_p(u, v, rr, norm))   # :142-143
    B = pc(_embed(sx, m, N), _embed(sy, m, N), r)
    A = pc(_embed(sx, m + 1, N), _embed(sy, m + 1, N), r)
    if A == 0:
        return np.inf
    return -np.log(A / B)


# ------------------------------------------------------------------ syntropy.mixed, knn branch
# (the caller of differential_entropy per discrete class.  Only
# ``continuous_estimator="knn"`` is on the path; the Gaussian branch belongs to syntropy.gaussian.)
def _mixed_classes(discrete_vars):
    """mixed/shannon.py:53-56,63-65: the distinct columns (np.unique, axis=1: lexicographic
    order), their probabilities and one boolean mask per class."""
    discrete_vars = np.asarray(discrete_vars)
    unq, counts = np.unique(discrete_vars, axis=1, return_counts=True)
    probs = counts / counts.sum()
    masks = [np.all(discrete_vars == unq[:, [i]], axis=0) for i in range(unq.shape[1])]
    return probs, masks


def mixed_shannon_entropy(discrete_vars, continuous_vars, k=5):
    """mixed/shannon.py:7-90 with continuous_estimator="knn": H(X,Y) = H(X) + H(Y|X)."""
    discrete_vars, continuous_vars = np.asarray(discrete_vars), np.asarray(continuous_vars)
    assert discrete_vars.shape[1] == continuous_vars.shape[1]
    probs, masks = _mixed_classes(discrete_vars)
    avg = -np.sum(probs * np.log(probs))
    ptw = np.zeros((1, continuous_vars.shape[1]))
    conditional = 0.0
    for i, mask in enumerate(masks):
        ptw[:, mask] -= np.log(probs[i])
        ptw_c, avg_c = differential_entropy(continuous_vars[:, mask], k=k)
        ptw[:, mask] += ptw_c
        conditional += probs[i] * avg_c
    avg += conditional
    return ptw, avg


def mixed_conditional_entropy(discrete_vars, continuous_vars, conditional="discrete", k=5):
    """mixed/shannon.py:93-181 with continuous_estimator="knn": joint minus the marginal of the
    conditioning variable."""
    discrete_vars, continuous_vars = np.asarray(discrete_vars), np.asarray(continuous_vars)
    assert discrete_vars.shape[1] == continuous_vars.shape[1]
    assert conditional in {"discrete", "continuous"}
    ptw_joint, avg_joint = mixed_shannon_entropy(discrete_vars, continuous_vars, k=k)
    if conditional == "continuous":
        ptw_marginal, avg_marginal = differential_entropy(continuous_vars, k)
    else:
        probs, masks = _mixed_classes(discrete_vars)
        ptw_marginal = np.zeros((1, discrete_vars.shape[1]))
        avg_marginal = -np.sum(probs * np.log(probs))
        for i, mask in enumerate(masks):
            ptw_marginal[:, mask] = -np.log(probs[i])
    return ptw_joint - ptw_marginal, avg_joint - avg_marginal


def mixed_mutual_information(discrete_vars, continuous_vars, k=5):
    """mixed/shannon.py:184-275 with continuous_estimator="knn": I(D;C) = H(C) - H(C|D)."""
    discrete_vars, continuous_vars = np.asarray(discrete_vars), np.asarray(continuous_vars)
    assert discrete_vars.shape[1] == continuous_vars.shape[1]
    ptw, avg = differential_entropy(continuous_vars, k=k)
    probs, masks = _mixed_classes(discrete_vars)
    conditional = 0.0
    for i, mask in enumerate(masks):
        local_c, avg_c = differential_entropy(continuous_vars[:, mask], k=k)
        conditional += probs[i] * avg_c
        ptw[:, mask] -= local_c
    avg -= conditional
    return ptw, avg
