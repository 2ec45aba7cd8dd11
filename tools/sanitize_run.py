#!/usr/bin/env python
"""Small-n run of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck /
initcheck): windowed + dense + exact engines, own-order and fused counts, leave-one-out counts,
KSG-2, KL divergence (cross-set search), pair counts, the Minkowski kernels, the sample sort and the
early-order prepare with its side streams.  Prints SANITIZE_RUN_OK when every result also equals the
CPU oracle.

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ksg_oracle as orc  # noqa: E402
from syntropy_b200 import _engine as E  # noqa: E402
from syntropy_b200 import knn, workloads  # noqa: E402


def same(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ok = np.allclose(a, b, rtol=1e-12, atol=1e-12, equal_nan=True)
    print(f"  {what}: {'ok' if ok else 'MISMATCH'}", flush=True)
    return ok


def main():
    n = int(os.environ.get("SANITIZE_N", "2500"))
    ok = True
    rng = np.random.default_rng(0)
    for engine in ("windowed", "dense", "exact"):
        E.config.engine = engine
        print(engine, flush=True)
        d, c = workloads.c1_readme_mi(n=n, seed=1)
        ok &= same(knn.mutual_information(c["idxs_x"], c["idxs_y"], c["k"], d)[0], orc.mutual_information(c["idxs_x"], c["idxs_y"], c["k"], d)[0], "mi 1D;1D")
        d3, c3 = workloads.c3_cmi(n=n, seed=2)
        ok &= same(knn.conditional_mutual_information(c3["idxs_x"], c3["idxs_y"], c3["idxs_z"], c3["k"], d3)[0],
                   orc.conditional_mutual_information(c3["idxs_x"], c3["idxs_y"], c3["idxs_z"], c3["k"], d3)[0], "cmi 2D;2D|2D")
        d4, c4 = workloads.c4_oinfo(n=n // 2, seed=3)
        ok &= same(knn.o_information(d4, c4["k"])[0], orc.o_information(d4, c4["k"])[0], "o-information 8 vars")
        ok &= same(knn.total_correlation(d4, c4["k"], algorithm=2)[0], orc.total_correlation(d4, c4["k"], algorithm=2)[0], "tc, KSG-2")
        d5, c5 = workloads.c5_info_rate(n=n, seed=4)
        ok &= same(knn.mutual_information(c5["idxs_x"], c5["idxs_y"], c5["k"], d5)[0],
                   orc.mutual_information(c5["idxs_x"], c5["idxs_y"], c5["k"], d5)[0], "mi 4D;4D")
        post, prior = rng.standard_normal((3, n)), 1.2 * rng.standard_normal((3, n + 300)) + 0.1
        ok &= same(knn.kullback_leibler_divergence(post, prior, 2)[0], orc.kullback_leibler_divergence(post, prior, 2)[0], "KL 3D")
        series = workloads.c5_series(n=n, seed=5)
        ok &= same(knn.sample_entropy((0,), series), orc.sample_entropy((0,), series), "sample entropy")
        ok &= same(knn.cross_sample_entropy((0,), (1,), series), orc.cross_sample_entropy((0,), (1,), series), "cross sample entropy")
        tied = np.round(rng.standard_normal((2, n)) * 6) / 6
        with np.errstate(all="ignore"):
            ok &= same(knn.mutual_information((0,), (1,), 3, tied)[0], orc.mutual_information((0,), (1,), 3, tied)[0], "tied data (fallbacks)")
    E.config.engine = "windowed"
    print("minkowski", flush=True)
    for p in (1, 2):
        ok &= same(knn.conditional_mutual_information((0, 1), (2,), (3, 4), 3, d3[:5, :1500], p=p)[0],
                   orc.conditional_mutual_information((0, 1), (2,), (3, 4), 3, d3[:5, :1500], p=p)[0], f"cmi p={p}")
        ok &= same(knn.sample_entropy((0,), series[:, :1500], norm=p), orc.sample_entropy((0,), series[:, :1500], norm=p), f"sampen p={p}")
    print("sample sort", flush=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_sort import gpu_sort

    keys = rng.standard_normal(20_000)
    got_k, got_i = gpu_sort(keys)
    ok &= bool(np.array_equal(got_i, np.argsort(keys, kind="stable")))
    torch.cuda.synchronize()
    print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_MISMATCH", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
