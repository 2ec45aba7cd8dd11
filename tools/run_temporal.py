#!/usr/bin/env python
"""knn.temporal on the C5 series (secondary): sample_entropy / cross_sample_entropy
timing on the GPU, and bit parity of the result against scipy's cKDTree.count_neighbors (the
reference's own call, knn/temporal.py:65-72) at a size the CPU finishes in about a minute.

    python tools/run_temporal.py [--n 4000000] [--check-n 400000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from syntropy_b200 import knn, workloads  # noqa: E402


def reference_sample_entropy(series, m=2, r=0.2):
    from numpy.lib.stride_tricks import sliding_window_view
    from scipy.spatial import cKDTree

    N = series.shape[0] - m
    small = cKDTree(sliding_window_view(series, m)[:N])
    big = cKDTree(sliding_window_view(series, m + 1))
    r = r * np.std(series)
    a = small.count_neighbors(small, r, p=np.inf) - N
    b = big.count_neighbors(big, r, p=np.inf) - N
    return -np.log(b / a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4_000_000)
    ap.add_argument("--check-n", type=int, default=400_000)
    args = ap.parse_args()
    out = {}
    if args.check_n:
        s = workloads.c5_series(n=args.check_n, seed=0)
        knn.sample_entropy((0,), s)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = knn.sample_entropy((0,), s)
        out["check"] = {"n": args.check_n, "gpu_ms": 1e3 * (time.perf_counter() - t0), "sampen": float(got)}
        t0 = time.perf_counter()
        want = reference_sample_entropy(s[0])
        out["check"]["cpu_s"] = time.perf_counter() - t0
        out["check"]["equal_to_reference"] = bool(got == want)
    s = workloads.c5_series(n=args.n, seed=0)
    knn.sample_entropy((0,), s[:, :100_000])
    for name, fn in (("sample_entropy_ch0", lambda: knn.sample_entropy((0,), s)),
                     ("sample_entropy_ch3", lambda: knn.sample_entropy((3,), s)),
                     ("cross_sample_entropy_01", lambda: knn.cross_sample_entropy((0,), (1,), s))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v = fn()
        torch.cuda.synchronize()
        out[name] = {"n": args.n, "ms": 1e3 * (time.perf_counter() - t0), "value": float(v)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
