#!/usr/bin/env python
"""CPU-only analysis (numpy + scipy, no GPU): what do the filter warps of the windowed k-NN kernel
spend their list maintenance on?  Emulates knn_plan_kernel + knn_units_kernel for a sample of
128-query blocks of the C2 generator (one query per lane, 32-row sub-tiles, seeds from +-24 rows,
unsorted replace-the-maximum lists of kcap entries) and counts, per query: inserts while the list
still has free slots vs replacements (each replacement = one sweep over the list), and per scanned
sub-tile: hit bits per lane vs the rounds the lock-step pop loop needs (= the busiest lane's bits).
Output of the round-1 run: profiles/r1aj_hit_simulation.txt.

    python tools/hit_sim.py [--n 200000] [--blocks 60]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from syntropy_b200 import workloads  # noqa: E402
from pruning_sim import prepared, seeds  # noqa: E402

TILE, SUB, WARP, BLOCK = 512, 32, 32, 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--blocks", type=int, default=60)
    args = ap.parse_args()
    data, call = workloads.c2_mi_1m(n=args.n, seed=0)
    k = call["k"]
    kcap = k + 1 + 3
    pts = prepared(data)
    n, dim = pts.shape
    n_tiles = (n + TILE - 1) // TILE
    pad = np.full((n_tiles * TILE - n, dim), np.inf, dtype=np.float32)
    rows = np.concatenate([pts, pad])
    tile_lo = rows.reshape(n_tiles, TILE, dim).min(axis=1)
    tile_hi = np.where(np.isinf(rows), -np.inf, rows).reshape(n_tiles, TILE, dim).max(axis=1)
    sub_lo = rows.reshape(-1, SUB, dim).min(axis=1)
    sub_hi = np.where(np.isinf(rows), -np.inf, rows).reshape(-1, SUB, dim).max(axis=1)
    seed = seeds(pts, kcap, 24)
    seed = np.nextafter(seed, np.float32(np.inf))
    rng = np.random.default_rng(1)
    blocks = rng.choice(n // BLOCK, size=args.blocks, replace=False)
    fill = repl = scans = bits_total = rounds_total = culled = 0
    stale = 0
    for b in blocks:
        q0 = b * BLOCK
        q = pts[q0:q0 + BLOCK]
        thr = seed[q0:q0 + BLOCK].copy()
        lists = np.full((BLOCK, kcap), np.inf, dtype=np.float32)
        nfill = np.zeros(BLOCK, dtype=np.int64)
        blo, bhi, R = q.min(axis=0), q.max(axis=0), thr.max()
        gap = np.maximum(np.maximum(tile_lo - bhi, blo - tile_hi), 0).max(axis=1)
        own = (q0 + BLOCK // 2) // TILE
        keep = np.flatnonzero(gap <= R)
        keep = keep[np.argsort(np.abs(keep - own), kind="stable")]  # near tiles first
        for t in keep:
            for w in range(BLOCK // WARP):
                sl = slice(w * WARP, (w + 1) * WARP)
                wq = q[sl]
                wlo, whi = wq.min(axis=0), wq.max(axis=0)
                for s in range(TILE // SUB):
                    sb = t * (TILE // SUB) + s
                    g = np.maximum(np.maximum(sub_lo[sb] - whi, wlo - sub_hi[sb]), 0).max()
                    if not g < thr[sl].max():
                        culled += 1
                        continue
                    scans += 1
                    cand = rows[sb * SUB:(sb + 1) * SUB]
                    d = np.abs(wq[:, None, :] - cand[None, :, :]).max(axis=2)  # (32 queries, 32 candidates)
                    mask = d < thr[sl][:, None]
                    bits = mask.sum(axis=1)
                    bits_total += int(bits.sum())
                    rounds_total += int(bits.max())
                    for lane in np.flatnonzero(bits):
                        qi = w * WARP + lane
                        for c in np.flatnonzero(mask[lane]):
                            m = d[lane, c]
                            if not m < thr[qi]:
                                stale += 1
                                continue
                            if nfill[qi] < kcap:
                                lists[qi, nfill[qi]] = m
                                nfill[qi] += 1
                                fill += 1
                                if nfill[qi] == kcap:
                                    thr[qi] = min(thr[qi], lists[qi].max())
                            else:
                                lists[qi, lists[qi].argmax()] = m
                                repl += 1
                                thr[qi] = min(thr[qi], lists[qi].max())
    nq = len(blocks) * BLOCK
    print(f"C2 generator, n = {n}, kcap = {kcap}, {len(blocks)} blocks of {BLOCK} queries")
    print(f"  sub-tile tests per query-warp: {(scans + culled) / (nq / WARP):.1f}, scanned: {scans / (nq / WARP):.1f} "
          f"({32 * scans / (nq / WARP):.0f} pairs per query)")
    print(f"  inserts per query: {(fill + repl) / nq:.2f}  = {fill / nq:.2f} appended while the list had free slots "
          f"+ {repl / nq:.2f} replacements (one sweep each); hit bits that were stale when popped: {stale / nq:.2f} per query")
    print(f"  hit bits per scanned sub-tile: {bits_total / scans:.1f} over 32 lanes; lock-step pop rounds: "
          f"{rounds_total / scans:.2f}  -> lanes busy in the pop loop: {100 * bits_total / (32 * max(rounds_total, 1)):.0f} %")


if __name__ == "__main__":
    main()
