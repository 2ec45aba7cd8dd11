#!/usr/bin/env python
"""Randomised agreement test: the windowed and dense fp32-filtered engines against the exact float64
engine, bit for bit, on radii and all subspace counts -- random sizes (not multiples of the tile /
sub-tile lengths), dimensions 1..8, k, subspace families, prepared orders, and awkward
distributions (clusters, exact duplicates, a constant coordinate, quantised values, heavy tails,
huge offsets).  Also KSG-2, entropy, KL divergence and pair counts through the public API.

    python tools/fuzz_engines.py [--cases 60] [--seed 0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from syntropy_b200 import _engine as E  # noqa: E402
from syntropy_b200 import knn  # noqa: E402


def make_data(rng, dim, n):
    kind = rng.choice(["normal", "uniform", "clusters", "duplicates", "constant_dim", "quantised", "heavy", "offset",
                       "correlated"])
    if kind == "normal":
        x = rng.standard_normal((dim, n))
    elif kind == "uniform":
        x = rng.uniform(-1, 1, (dim, n))
    elif kind == "clusters":
        centres = rng.standard_normal((dim, 5)) * 10
        x = centres[:, rng.integers(0, 5, n)] + 0.05 * rng.standard_normal((dim, n))
    elif kind == "duplicates":
        x = rng.standard_normal((dim, n))
        x[:, n // 2:] = x[:, : n - n // 2]            # every point twice
    elif kind == "constant_dim":
        x = rng.standard_normal((dim, n))
        x[rng.integers(0, dim)] = 3.25
    elif kind == "quantised":
        x = np.round(rng.standard_normal((dim, n)) * 8) / 8
    elif kind == "heavy":
        x = rng.standard_t(1.3, (dim, n))
    elif kind == "offset":
        x = rng.standard_normal((dim, n)) * 1e-3 + 1e6
    else:
        base = rng.standard_normal((1, n))
        x = 0.8 * base + 0.3 * rng.standard_normal((dim, n))
    return kind, x


def subspaces(rng, dim):
    subs = [(d,) for d in range(dim)]
    if dim >= 2:
        subs.append(tuple(range(dim)))
        cut = int(rng.integers(1, dim))
        subs += [tuple(range(cut)), tuple(range(cut, dim))]
    if dim >= 3:
        subs += [tuple(j for j in range(dim) if j != d) for d in range(dim)]
        pick = sorted(rng.choice(dim, size=int(rng.integers(2, dim)), replace=False).tolist())
        subs.append(tuple(pick))
    return list(dict.fromkeys(subs))


def radius_and_counts(rows, k, subs):
    pts = E.to_device_points(rows)
    masks = [E.mask_of(s) for s in subs]
    n = rows.shape[1]
    if not E._use_fast(rows.shape[0], k + 1):
        eps = E.knn_radius(pts, k + 1, 0, n)
        counts = E.count_subspaces(pts, masks, eps, 0, n)
    else:
        eps, counts, _ = E.radius_and_counts(pts, k, masks)
    torch.cuda.synchronize()
    return eps.cpu().numpy(), counts.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--big", action="store_true", help="sizes 5e4 .. 3e5 instead of 37 .. 2e4")
    ap.add_argument("--only", type=int, default=-1, help="evaluate this case only (the others just advance the generator)")
    ap.add_argument("--repeat", type=int, default=1, help="evaluate every selected case this many times")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = 0
    for case in range(args.cases):
        dim = int(rng.integers(1, 9))
        n = int(rng.choice([50_021, 131_072, 200_003, 300_007] if args.big else
                           [37, 200, 511, 513, 1000, 2049, 5000, 9973, 20011]))
        k = int(rng.integers(1, min(12, n - 1)))
        kind, rows = make_data(rng, dim, n)
        subs = subspaces(rng, dim)
        if args.only >= 0 and case != args.only:
            continue
        for _rep in range(args.repeat):
            E.config.engine, E.config.primary = "exact", -1
            ref = radius_and_counts(rows, k, subs)
            for engine, primary in (("windowed", -1), ("windowed", -2), ("dense", -1)):
                E.config.engine, E.config.primary = engine, primary
                got = radius_and_counts(rows, k, subs)
                ok = np.array_equal(got[0], ref[0], equal_nan=True) and np.array_equal(got[1], ref[1])
                if not ok:
                    bad += 1
                    print(f"MISMATCH case {case}: {kind} dim={dim} n={n} k={k} engine={engine} primary={primary} "
                          f"eps_diff={int((got[0] != ref[0]).sum())} count_diff={int((got[1] != ref[1]).sum())}", flush=True)
            # public API under the production engine vs the exact engine
            api = {}
            for engine in ("exact", "windowed"):
                E.config.engine, E.config.primary = engine, -1
                out = [knn.differential_entropy(rows, k)]
                if dim >= 2:
                    cut = dim // 2
                    x, y = tuple(range(cut)), tuple(range(cut, dim))
                    out.append(knn.mutual_information(x, y, k, rows))
                    if kind not in ("duplicates", "quantised", "constant_dim"):   # KSG-2: tie-free data only
                        out.append(knn.mutual_information(x, y, k, rows, algorithm=2))
                    other = rows[:, ::-1] * 1.1 + 0.05
                    out.append(knn.kullback_leibler_divergence(rows, np.ascontiguousarray(other[:, : max(k + 2, n // 2)]), k))
                if n >= 200:
                    out.append((np.array([knn.sample_entropy((0,), rows, m=2, r=0.3)]), 0.0))
                api[engine] = out
            for which, (a, b) in enumerate(zip(api["exact"], api["windowed"])):
                with np.errstate(all="ignore"):
                    same = np.array_equal(np.asarray(a[0]), np.asarray(b[0]), equal_nan=True)
                    close = np.allclose(np.asarray(a[0]), np.asarray(b[0]), rtol=1e-12, atol=1e-12, equal_nan=True)
                if not (same or close):
                    bad += 1
                    aa, bb = np.asarray(a[0], dtype=np.float64).ravel(), np.asarray(b[0], dtype=np.float64).ravel()
                    with np.errstate(all="ignore"):
                        worst = int(np.nanargmax(np.abs(aa - bb))) if aa.shape == bb.shape and aa.size else -1
                    ex = repr(aa[worst]) if worst >= 0 else "?"
                    wi = repr(bb[worst]) if worst >= 0 else "?"
                    nd = int((aa != bb).sum()) if aa.shape == bb.shape else -1
                    print(f"API MISMATCH case {case}: {kind} dim={dim} n={n} k={k} output#{which} n_diff={nd} "
                          f"worst_at={worst} exact={ex} windowed={wi} avg_exact={a[1]!r} avg_windowed={b[1]!r}", flush=True)
            print(f"case {case}: {kind} dim={dim} n={n} k={k} subs={len(subs)} ok", flush=True)
    print("FUZZ_OK" if bad == 0 else f"FUZZ_FAILED {bad}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
