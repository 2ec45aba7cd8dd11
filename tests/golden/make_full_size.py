"""Generate tests/golden/full_size.json (+ full_size_c5.npz): the BASELINE configurations at
their FULL sizes, computed on the CPU with the "fast oracle" of SURVEY.md 8(c).

Run from the repo root:  ``python tests/golden/make_full_size.py [c2 c3 c4 c5]``

The estimators are the ones of ``oracle/ksg_oracle.py`` (pinned on the reference's known
answers and on fixtures produced by the unmodified reference, tests/test_oracle.py) with its
two brute-force primitives swapped for the reference's own accelerator, ``scipy.spatial.
cKDTree`` -- ``tree.query(.., k + 1, p=inf, workers=-1)`` and one vectorised
``tree.query_ball_point(X, nextafter(eps, -inf), p=inf, return_length=True, workers=-1) - 1``,
which SURVEY.md 8(c) found ``array_equal`` to the reference's per-sample Python loop.  The
float64 operation order of every pointwise array is therefore the reference's
(knn/shannon.py:188-196,329-347; knn/multivariate_mi.py:100-115,241-260,311-335,394-415).

What is stored per configuration: n, k, the average, mean |ptw| (the floor of the 1e-12
relative tolerance), Neumaier-compensated sum, and the pointwise values / radii at 64 fixed
sample indices.  C5 (n = 4e6, 8-D) is too slow for a full CPU pass: a fixed 50 000-query
subset is evaluated against the FULL point set; radii, both marginal counts and the pointwise
values of that subset go to full_size_c5.npz.

Provenance of the committed numbers: scipy 1.18.1, numpy 2.3.5, Python 3.12.3 (x86-64).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np
from scipy.spatial import cKDTree

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ksg_oracle as orc  # noqa: E402
from syntropy_b200 import workloads  # noqa: E402

_trees: dict = {}
_radii: dict = {}
_counts: dict = {}


def _key(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return hashlib.sha1(a.tobytes()).hexdigest(), a


def _tree(points):
    key, pts = _key(points)
    if key not in _trees:
        _trees[key] = cKDTree(pts.T)
    return key, pts, _trees[key]


def fast_radius(points, k, queries=None, block=256, return_indices=False):
    assert not return_indices
    key, pts, tree = _tree(points)
    if queries is None:
        if (key, k) not in _radii:
            dist, _ = tree.query(pts.T, k=k + 1, p=np.inf, workers=-1)
            _radii[(key, k)] = np.ascontiguousarray(dist[:, -1])
        return _radii[(key, k)]
    qs = np.ascontiguousarray(queries, dtype=np.float64)
    dist, _ = tree.query(qs.T, k=k, p=np.inf, workers=-1)
    return np.ascontiguousarray(dist[:, -1] if k > 1 else dist.reshape(-1))


def fast_count(points, eps, strict=True, queries=None, block=256):
    key, pts, tree = _tree(points)
    eps = np.asarray(eps, dtype=np.float64)
    r = np.nextafter(eps, -np.inf) if strict else eps
    if queries is None:
        ck = (key, hashlib.sha1(r.tobytes()).hexdigest())
        if ck not in _counts:
            _counts[ck] = tree.query_ball_point(pts.T, r, p=np.inf, return_length=True, workers=-1).astype(np.int64) - 1
        return _counts[ck]
    qs = np.ascontiguousarray(queries, dtype=np.float64)
    return tree.query_ball_point(qs.T, r, p=np.inf, return_length=True, workers=-1).astype(np.int64) - 1


orc.knn_radius = fast_radius
orc.count_within = fast_count

PROBE = 64


def neumaier(values):
    s = 0.0
    c = 0.0
    for v in values:
        t = s + v
        c += (s - t) + v if abs(s) >= abs(v) else (v - t) + s
        s = t
    return s + c


def summary(ptw, avg, n):
    flat = np.asarray(ptw, dtype=np.float64).reshape(-1)
    idx = np.sort(np.random.default_rng(7).choice(n, size=PROBE, replace=False))
    finite = flat[np.isfinite(flat)]
    import math

    return {"avg": float(avg), "mean_abs_ptw": float(np.mean(np.abs(finite))), "fsum_over_n": math.fsum(flat) / n,
            "probe_idx": [int(i) for i in idx], "probe_ptw": [float(v) for v in flat[idx]]}


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c2", "c3", "c4", "c5"]
    path = os.path.join(HERE, "full_size.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    out["provenance"] = "scipy cKDTree fast oracle through oracle/ksg_oracle.py estimators; see make_full_size.py"
    for name in which:
        t0 = time.time()
        if name == "c2":
            data, call = workloads.c2_mi_1m()
            ptw, avg = orc.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)
            out["c2"] = {"n": data.shape[1], "k": call["k"], "mi": summary(ptw, avg, data.shape[1])}
        elif name == "c3":
            data, call = workloads.c3_cmi()
            ptw, avg = orc.conditional_mutual_information(call["idxs_x"], call["idxs_y"], call["idxs_z"], call["k"], data)
            out["c3"] = {"n": data.shape[1], "k": call["k"], "cmi": summary(ptw, avg, data.shape[1])}
        elif name == "c4":
            data, call = workloads.c4_oinfo()
            n = data.shape[1]
            rec = {"n": n, "k": call["k"]}
            for fn in ("o_information", "total_correlation", "dual_total_correlation", "s_information"):
                ptw, avg = getattr(orc, fn)(data, call["k"])
                rec[fn] = summary(ptw, avg, n)
                print("  ", fn, rec[fn]["avg"], f"{time.time() - t0:.0f}s", flush=True)
            out["c4"] = rec
        elif name == "c5":
            data, call = workloads.c5_info_rate()
            n, k = data.shape[1], call["k"]
            sel = np.sort(np.random.default_rng(11).choice(n, size=50_000, replace=False))
            joint = data[call["idxs_x"] + call["idxs_y"], :]
            eps = fast_radius(joint, k + 1, queries=joint[:, sel])  # k' = k + 1: the query is a member of the set
            psi_k, psi_n = float(orc.psi_int(k)), float(orc.psi_int(n))
            ptw = np.full(sel.size, psi_k + psi_n)
            counts = []
            for idxs in (call["idxs_x"], call["idxs_y"]):
                rows = data[idxs, :]
                c = fast_count(rows, eps, queries=rows[:, sel])
                counts.append(c.astype(np.int32))
                ptw -= orc.psi_int(c + 1)
            np.savez_compressed(os.path.join(HERE, "full_size_c5.npz"), sel=sel.astype(np.int32), eps=eps,
                                counts=np.stack(counts), ptw=ptw)
            out["c5"] = {"n": n, "k": k, "subset": int(sel.size), "subset_seed": 11,
                         "subset_mean": float(ptw.mean()), "mean_abs_ptw": float(np.mean(np.abs(ptw)))}
        print(name, json.dumps({k_: v for k_, v in out[name].items() if not isinstance(v, dict)}),
              f"{time.time() - t0:.0f}s", flush=True)
        json.dump(out, open(path, "w"), indent=1)
        _trees.clear(); _radii.clear(); _counts.clear()


if __name__ == "__main__":
    main()
