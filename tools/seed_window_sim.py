#!/usr/bin/env python
"""CPU-only analysis (numpy + scipy, no GPU): how tight is the planning pass's seed radius -- the
kcap-th smallest distance among the +-W neighbours of a query on the Hilbert curve -- compared with
the true kcap-th neighbour distance, as a function of the window width W?  (ratio)^dim is the factor
by which the seeded ball's volume exceeds the final one.  Run for the C3 / C5 / C2 generators at
n = 2e5; output of the round-1 run: profiles/r1aj_seed_window_simulation.txt.

    python tools/seed_window_sim.py
"""
import sys, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from syntropy_b200 import workloads
from scipy.spatial import cKDTree

def hilbert_keys(ranks, bits):
    # ranks: (dim, n) uint32 with `bits` significant bits each; Skilling's axes-to-transpose, vectorised
    X = ranks.astype(np.uint64).copy()
    dim, n = X.shape
    M = np.uint64(1) << np.uint64(bits-1)
    Q = M
    while Q > 1:
        P = Q - np.uint64(1)
        for d in range(dim):
            hit = (X[d] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = (X[0] ^ X[d]) & P
            t = np.where(hit, np.uint64(0), t)
            X[0] ^= t; X[d] ^= t
        Q >>= np.uint64(1)
    for d in range(1, dim): X[d] ^= X[d-1]
    t = np.zeros(n, dtype=np.uint64)
    Q = M
    while Q > 1:
        t = np.where((X[dim-1] & Q) != 0, t ^ (Q - np.uint64(1)), t)
        Q >>= np.uint64(1)
    for d in range(dim): X[d] ^= t
    key = np.zeros(n, dtype=object)  # python ints: up to dim*bits bits
    keys = [0]*n
    # build as two uint64 halves to sort lexicographically
    hi = np.zeros(n, dtype=np.uint64); lo = np.zeros(n, dtype=np.uint64); nb = 0
    total = bits*dim
    for level in range(bits):
        for d in range(dim):
            bit = (X[d] >> np.uint64(bits-1-level)) & np.uint64(1)
            pos = total - 1 - nb
            if pos >= 64: hi |= bit << np.uint64(pos-64)
            else: lo |= bit << np.uint64(pos)
            nb += 1
    return hi, lo

def run(name, data, k, n):
    dim = data.shape[0]
    ranks = np.argsort(np.argsort(data, axis=1, kind='stable'), axis=1).astype(np.uint32)
    rb = int(np.ceil(np.log2(n)))
    bits = min(63//dim, max(rb-4,1))
    hi, lo = hilbert_keys(ranks >> np.uint32(rb-bits), bits)
    order = np.lexsort((lo, hi))
    pts = data[:, order].T.astype(np.float32)
    kcap = k + 1 + 3
    tree = cKDTree(pts)
    dist, _ = tree.query(pts, k=kcap, p=np.inf)
    true_kcap = dist[:, -1]
    rng = np.random.default_rng(0)
    sample = rng.choice(n, size=4000, replace=False)
    print(f"{name}: n={n} dim={dim} kcap={kcap}")
    for W in (24, 48, 96, 192, 384):
        ratios = []
        for i in sample:
            lo_i, hi_i = max(0, i-W), min(n, i+W+1)
            d = np.abs(pts[lo_i:hi_i] - pts[i]).max(axis=1)
            d = np.delete(d, i-lo_i)
            # self counts as one candidate in the real kernel? the kernel skips self (self_skip) and kcap includes self
            seed = np.sort(np.concatenate([[0.0], d]))[kcap-1]
            ratios.append(seed / true_kcap[i])
        r = np.array(ratios)
        print(f"  W={W:4d}: median seed/true = {np.median(r):.3f}, mean volume factor (ratio^dim) = {np.mean(r**dim):.1f}, median {np.median(r**dim):.1f}")

def main():
    n = 200_000
    d3, c3 = workloads.c3_cmi(n=n, seed=0)
    run("C3 joint 6-D", d3, c3["k"], n)
    d5, c5 = workloads.c5_info_rate(n=n, seed=0)
    run("C5 joint 8-D", d5, c5["k"], n)
    d2, c2 = workloads.c2_mi_1m(n=n, seed=0)
    run("C2 joint 2-D", d2, c2["k"], n)


if __name__ == "__main__":
    main()
