#!/usr/bin/env python
"""Times the knn branch of syntropy_b200.mixed on a synthetic discrete x
continuous sample and checks a random subset of every class against the CPU oracle.

    python tools/run_mixed.py [--n 1000000] [--classes 4] [--dim 2] [--k 5] [--reps 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from syntropy_b200 import mixed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--classes", type=int, default=4)
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=64, help="queries per class checked against the oracle")
    args = ap.parse_args()
    from oracle import ksg_oracle_c as oc

    rng = np.random.default_rng(0)
    d = rng.integers(0, args.classes, size=(1, args.n))
    c = rng.standard_normal((args.dim, args.n)) * (1.0 + 0.25 * d) + 1.5 * d
    out = {"n": args.n, "classes": args.classes, "dim": args.dim, "k": args.k}
    for name, fn in (("mutual_information", lambda: mixed.mutual_information(d, c, "knn", args.k)),
                     ("shannon_entropy", lambda: mixed.shannon_entropy(d, c, "knn", args.k))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            ptw, avg = fn()
        torch.cuda.synchronize()
        out[name + "_ms"] = 1e3 * (time.perf_counter() - t0) / args.reps
        out[name + "_avg"] = float(avg)
    # parity of the per-class radii behind shannon_entropy: ptw = -log p + (-psi(k) + psi(n_c) + D log(2 eps))
    ptw, _ = mixed.shannon_entropy(d, c, "knn", args.k)
    from scipy.special import digamma

    ok = True
    for state in range(args.classes):
        idx = np.flatnonzero(d[0] == state)
        sub = np.ascontiguousarray(c[:, idx])
        pick = rng.choice(idx.shape[0], size=min(args.check, idx.shape[0]), replace=False)
        eps = oc.knn_radius(sub, args.k + 1, queries=np.ascontiguousarray(sub[:, pick]))
        p = idx.shape[0] / args.n
        want = (0.0 - np.log(p)) + (-digamma(args.k) + digamma(idx.shape[0]) + args.dim * np.log(2 * eps))
        got = ptw[0, idx[pick]]
        ok = ok and bool(np.all(np.abs(got - want) <= 1e-12 * np.maximum(np.abs(want), 1.0)))
    out["oracle_subset_parity"] = ok
    print(json.dumps(out))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
