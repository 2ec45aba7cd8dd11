#!/usr/bin/env python
"""CPU-only analysis (numpy + scipy, no GPU): which fraction of all n^2 (query, candidate) pairs does
sub-tile culling leave to be evaluated, for groups of G consecutive queries (the filter warp: box of
the group + its largest threshold) against sub-tiles of S consecutive rows, with the thresholds
fixed at (a) the true kcap-th neighbour distance -- the lower bound of the work for that geometry --
and (b) the planning pass's seed for several window widths?  Output of the round-1 run:
profiles/r1aj_pruning_simulation.txt.

    python tools/pruning_sim.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from syntropy_b200 import workloads
from scipy.spatial import cKDTree
from seed_window_sim import hilbert_keys

def prepared(data):
    dim, n = data.shape
    ranks = np.argsort(np.argsort(data, axis=1, kind='stable'), axis=1).astype(np.uint32)
    rb = int(np.ceil(np.log2(n))); bits = min(63//dim, max(rb-4,1))
    hi, lo = hilbert_keys(ranks >> np.uint32(rb-bits), bits)
    order = np.lexsort((lo, hi))
    return data[:, order].T.astype(np.float32)

def seeds(pts, kcap, W):
    n = pts.shape[0]
    best = np.full((n, kcap), np.inf, dtype=np.float32); best[:,0] = 0.0   # self
    for off in range(1, W+1):
        for sgn in (-1, 1):
            j = np.arange(n) + sgn*off
            ok = (j >= 0) & (j < n)
            d = np.full(n, np.inf, dtype=np.float32)
            d[ok] = np.abs(pts[ok] - pts[j[ok]]).max(axis=1)
            best = np.sort(np.concatenate([best, d[:,None]], axis=1), axis=1)[:, :kcap]
    return best[:, -1]

def scanned_fraction(pts, thr, qgroup=32, sub=32, sample=400, seed=0):
    n, dim = pts.shape
    ns = n // sub
    sb = pts[:ns*sub].reshape(ns, sub, dim); slo = sb.min(axis=1); shi = sb.max(axis=1)
    ng = n // qgroup
    rng = np.random.default_rng(seed)
    gs = rng.choice(ng, size=min(sample, ng), replace=False)
    tot = 0
    for g in gs:
        q = pts[g*qgroup:(g+1)*qgroup]; wlo = q.min(axis=0); whi = q.max(axis=0)
        t = thr[g*qgroup:(g+1)*qgroup].max()
        gap = np.maximum(np.maximum(slo - whi, wlo - shi), 0).max(axis=1)
        tot += int((gap < t).sum())
    return tot / (len(gs) * ns)     # fraction of sub-tiles scanned = fraction of all pairs evaluated

def run(name, data, k, widths=(24, 96, 384)):
    n = data.shape[1]; dim = data.shape[0]
    pts = prepared(data)
    kcap = k + 1 + 3
    tree = cKDTree(pts)
    dist, _ = tree.query(pts, k=kcap, p=np.inf)
    true = dist[:, -1].astype(np.float32)
    print(f"{name}: n={n} dim={dim} kcap={kcap}; fraction of all n^2 pairs evaluated (sub-tile scans), static thresholds")
    for label, thr in [("true kcap-th radius (lower bound of the work)", true)] + [(f"seed, window +-{W}", seeds(pts, kcap, W)) for W in widths]:
        f32 = scanned_fraction(pts, thr, 32, 32)
        f16q = scanned_fraction(pts, thr, 16, 32)
        f16s = scanned_fraction(pts, thr, 32, 16)
        f8q = scanned_fraction(pts, thr, 8, 32)
        print(f"  {label:48s}: 32q x 32rows {100*f32:6.3f} %   16q x 32rows {100*f16q:6.3f} %   32q x 16rows {100*f16s:6.3f} %   8q x 32rows {100*f8q:6.3f} %")

def main():
    n = 200_000
    d3, c3 = workloads.c3_cmi(n=n, seed=0); run("C3 joint 6-D", d3, c3["k"])
    d5, c5 = workloads.c5_info_rate(n=n, seed=0); run("C5 joint 8-D (Hilbert order)", d5, c5["k"])
    d2, c2 = workloads.c2_mi_1m(n=n, seed=0); run("C2 joint 2-D", d2, c2["k"], widths=(24,))


if __name__ == "__main__":
    main()
