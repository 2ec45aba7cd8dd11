#!/usr/bin/env python
"""kNN KL divergence at the size of the reference's own test (tests/test_knn.py:78-113: two 3-D
Gaussian clouds of 1e6 samples, k = 1, checked against the closed form at 10 % relative), timing
on the GPU and, on a smaller sample, bit parity with the scipy call sequence of the reference
(knn/shannon.py:401-407).

    python tools/run_kl.py [--n 1000000] [--check-n 100000]
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

from syntropy_b200 import knn  # noqa: E402


def gaussian_kl(mu0, cov0, mu1, cov1):
    d = mu0.size
    inv1 = np.linalg.inv(cov1)
    diff = mu1 - mu0
    return 0.5 * (np.trace(inv1 @ cov0) + diff @ inv1 @ diff - d + np.log(np.linalg.det(cov1) / np.linalg.det(cov0)))


def reference_kl(post, prior, k):
    from scipy.spatial import cKDTree

    n, m, d = post.shape[1], prior.shape[1], post.shape[0]
    rho = cKDTree(post.T).query(post.T, k=k + 1, p=np.inf)[0][:, -1]
    nu = cKDTree(prior.T).query(post.T, k=k, p=np.inf)[0]
    nu = nu if nu.ndim == 1 else nu[:, -1]
    ptw = d * np.log(nu / rho) + np.log(m / (n - 1))
    return ptw, ptw.mean()


def clouds(n, seed=0):
    rng = np.random.default_rng(seed)
    mu0, mu1 = np.zeros(3), np.array([0.5, -0.25, 0.1])
    a = rng.standard_normal((3, 3)); cov0 = a @ a.T / 3 + np.eye(3)
    b = rng.standard_normal((3, 3)); cov1 = b @ b.T / 3 + np.eye(3)
    post = rng.multivariate_normal(mu0, cov0, size=n).T
    prior = rng.multivariate_normal(mu1, cov1, size=n).T
    return np.ascontiguousarray(post), np.ascontiguousarray(prior), gaussian_kl(mu0, cov0, mu1, cov1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--check-n", type=int, default=100_000)
    args = ap.parse_args()
    out = {}
    post, prior, _ = clouds(args.check_n, seed=1)
    knn.kullback_leibler_divergence(post, prior, 1)  # warm-up
    ptw, avg = knn.kullback_leibler_divergence(post, prior, 1)
    t0 = time.perf_counter()
    ref_ptw, ref_avg = reference_kl(post, prior, 1)
    out["check"] = {"n": args.check_n, "cpu_s": time.perf_counter() - t0,
                    "ptw_bit_equal": bool(np.array_equal(ptw, ref_ptw)),
                    "ptw_max_rel": float(np.max(np.abs(ptw - ref_ptw) / np.maximum(np.abs(ref_ptw), 1.0))),
                    "avg_rel": float(abs(avg - ref_avg) / abs(ref_avg))}
    post, prior, exact = clouds(args.n, seed=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ptw, avg = knn.kullback_leibler_divergence(post, prior, 1)
    torch.cuda.synchronize()
    out["full"] = {"n": args.n, "ms": 1e3 * (time.perf_counter() - t0), "estimate": float(avg),
                   "closed_form": float(exact), "rel_err": float(abs(avg - exact) / exact)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
