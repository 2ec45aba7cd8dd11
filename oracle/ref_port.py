"""CPU baseline port of the reference's KSG-1 path on scipy's cKDTree -- TEST/BENCH
INFRASTRUCTURE ONLY (used by ``bench.py``'s ``cpu_baseline`` and ``--impl reference``
legs and by tests; never imported by ``syntropy_b200``).

``/root/reference`` (pure Python) cannot travel to the GPU box, but its only native
dependency, scipy, is part of the image there.  This module issues the SAME scipy
calls in the SAME order as the reference's ``mutual_information_1``
(syntropy/knn/shannon.py:137-196 -> syntropy/knn/utils.py:6-35, 38-86):

    joint   : cKDTree(data[xy].T); tree.query(.., k=k+1, p=inf)        utils.py:29,33
    marginal: cKDTree(data[idx].T) + a (discarded) k-NN query          shannon.py:190
              nextafter(eps, -inf)                                     utils.py:76
              one tree.query_ball_point(x_i, eps_i, return_length=True) - 1 per sample,
              from a Python loop                                       utils.py:79-84
    psi     : scipy.special.digamma(counts + 1)                        shannon.py:194

``mode="faithful"`` reproduces that, single-threaded, Python loop included (what a
user of the reference gets: 1 core).  ``mode="fast"`` uses the same trees but one
vectorised ``query_ball_point`` over all samples and ``workers=-1`` -- bit-identical
results (SURVEY.md 8c "fast oracle"), and the strongest CPU arm the reference's own
dependency offers; it skips the discarded marginal k-NN queries.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.spatial import cKDTree
from scipy.special import digamma


def _counts(tree, rows, eps, mode, workers):
    eps_ = np.nextafter(eps, -np.inf)
    if mode == "faithful":
        return np.array([
            tree.query_ball_point(x_i, e_i, return_length=True, p=np.inf) - 1
            for x_i, e_i in zip(rows.T, eps_)
        ])
    return tree.query_ball_point(rows.T, eps_, p=np.inf, return_length=True, workers=workers) - 1


def mutual_information_1(idxs_x, idxs_y, k, data, mode="faithful"):
    assert mode in ("faithful", "fast")
    workers = 1 if mode == "faithful" else -1
    data = np.asarray(data)
    idxs_x, idxs_y = tuple(idxs_x), tuple(idxs_y)
    N = data.shape[1]
    psi_k, psi_N = digamma(k), digamma(N)
    joint = data[idxs_x + idxs_y, :]
    dist, _ = cKDTree(joint.T).query(joint.T, k=k + 1, p=np.inf, workers=workers)
    eps = dist[:, -1]
    ptw = np.full((1, N), psi_k + psi_N)
    for idxs in (idxs_x, idxs_y):
        rows = data[idxs, :]
        tree = cKDTree(rows.T)
        if mode == "faithful":
            tree.query(rows.T, k=k + 1, p=np.inf)  # computed and discarded by the reference
        ptw[0, :] -= digamma(_counts(tree, rows, eps, mode, workers) + 1)
    return ptw, ptw.mean()


def cores_used(mode: str) -> int:
    return 1 if mode == "faithful" else (os.cpu_count() or 1)
