"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run from the repo root:  ``python tests/golden/make_golden.py``

The reference (``/root/reference/syntropy``) is imported through a stub parent
package because ``import syntropy`` itself needs the installed distribution metadata
(``syntropy/__init__.py:11``); ``syntropy.knn`` and its relative imports resolve fine
under the stub.  ``/root/reference`` does not exist on the GPU box,
so only the small arrays written here travel.  Each fixture stores the *outputs* of
reference functions; the inputs are regenerated from ``syntropy_b200.workloads`` /
the seeds recorded in the file.

Provenance of the numbers in the committed files: scipy 1.18.1, numpy 2.3.5,
Python 3.12.3 (x86-64), reference checkout mounted at /root/reference.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from syntropy_b200 import workloads  # noqa: E402

REF = "/root/reference/syntropy"


def load_reference_knn():
    pkg = types.ModuleType("syntropy")
    pkg.__path__ = [REF]
    sys.modules["syntropy"] = pkg
    knn = importlib.import_module("syntropy.knn")
    utils = importlib.import_module("syntropy.knn.utils")
    return knn, utils


def radius_and_counts(utils, data, joint, subspaces, k):
    """eps from the reference's joint query and the strict counts of every subspace."""
    _, dist, _ = utils.build_tree_and_get_distances(data[tuple(joint), :], k=k)
    eps = dist[:, -1]
    out = {"eps": eps}
    for name, idxs in subspaces.items():
        sub = data[tuple(idxs), :]
        tree, _, _ = utils.build_tree_and_get_distances(sub, k=k)
        out["counts_" + name] = utils.get_counts_from_tree(tree, sub, eps)
    return out


def make_mixed():
    """tests/golden/mixed_knn.npz: the reference's syntropy.mixed functions with
    continuous_estimator="knn" on ``workloads.mixed_cases()``."""
    load_reference_knn()
    mixed = importlib.import_module("syntropy.mixed.shannon")
    out = {}
    with np.errstate(all="ignore"):
        for name, (d, c, k) in workloads.mixed_cases().items():
            ptw, avg = mixed.shannon_entropy(d, c, continuous_estimator="knn", k=k)
            out[f"{name}_h_ptw"], out[f"{name}_h_avg"] = ptw, avg
            for cond in ("discrete", "continuous"):
                ptw, avg = mixed.conditional_entropy(d, c, conditional=cond, continuous_estimator="knn", k=k)
                out[f"{name}_ce_{cond}_ptw"], out[f"{name}_ce_{cond}_avg"] = ptw, avg
            ptw, avg = mixed.mutual_information(d, c, continuous_estimator="knn", k=k)
            out[f"{name}_mi_ptw"], out[f"{name}_mi_avg"] = ptw, avg
    np.savez_compressed(os.path.join(HERE, "mixed_knn.npz"), **out)
    print("mixed_knn.npz", os.path.getsize(os.path.join(HERE, "mixed_knn.npz")))


def main():
    if "--only-mixed" in sys.argv:
        make_mixed()
        return
    knn, utils = load_reference_knn()
    os.makedirs(HERE, exist_ok=True)

    # ---- the reference's own tie-saturated integer fixture (tests/test_knn.py:20-24)
    tie = workloads.tie_fixture()
    out = {}
    for k in (1, 4):
        ptw, avg = knn.differential_entropy(data=tie, idxs=(0,), k=k)
        out[f"entropy_k{k}_ptw"], out[f"entropy_k{k}_avg"] = ptw, avg
    for alg in (1, 2):
        ptw, avg = knn.mutual_information((0,), (1,), k=1, data=tie, algorithm=alg)
        out[f"mi{alg}_ptw"], out[f"mi{alg}_avg"] = ptw, avg
    ptw, avg = knn.conditional_mutual_information((0,), (1,), (2,), k=1, data=tie)
    out["cmi_ptw"], out["cmi_avg"] = ptw, avg
    for name in ("total_correlation", "dual_total_correlation", "o_information", "s_information"):
        ptw, avg = getattr(knn, name)(data=tie, k=1)
        out[name + "_ptw"], out[name + "_avg"] = ptw, avg
    ptw, avg = knn.total_correlation(data=tie, k=1, algorithm=2)
    out["total_correlation2_ptw"], out["total_correlation2_avg"] = ptw, avg
    out.update({"rc_" + key: v for key, v in radius_and_counts(
        utils, tie, (0, 1, 2), {"x": (0,), "y": (1,), "z": (2,), "xy": (0, 1), "yz": (1, 2)}, 1).items()})
    np.savez_compressed(os.path.join(HERE, "tie_fixture.npz"), **out)

    # ---- C1 exactly (n=5000, k=5) and small versions of C2..C5
    data, call = workloads.c1_readme_mi(n=5000, seed=0)
    ptw, avg = knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)
    out = {"n": 5000, "seed": 0, "k": 5, "ptw": ptw, "avg": avg}
    out.update(radius_and_counts(utils, data, (0, 1), {"x": (0,), "y": (1,)}, 5))
    ptw2, avg2 = knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data, algorithm=2)
    out["ptw_alg2"], out["avg_alg2"] = ptw2, avg2
    np.savez_compressed(os.path.join(HERE, "c1_mi.npz"), **out)

    n = 3000
    data, call = workloads.c3_cmi(n=n, seed=0)
    ptw, avg = knn.conditional_mutual_information(call["idxs_x"], call["idxs_y"], call["idxs_z"], call["k"], data)
    out = {"n": n, "seed": 0, "k": 4, "ptw": ptw, "avg": avg}
    out.update(radius_and_counts(utils, data, (0, 1, 2, 3, 4, 5),
                                 {"z": (4, 5), "xz": (0, 1, 4, 5), "yz": (2, 3, 4, 5)}, 4))
    np.savez_compressed(os.path.join(HERE, "c3_cmi.npz"), **out)

    n = 2000
    data, call = workloads.c4_oinfo(n=n, seed=0)
    out = {"n": n, "seed": 0, "k": 4}
    for name in ("total_correlation", "dual_total_correlation", "o_information", "s_information"):
        ptw, avg = getattr(knn, name)(data=data, k=4)
        out[name + "_ptw"], out[name + "_avg"] = ptw, avg
    ptw, avg = knn.total_correlation(data=data, k=4, algorithm=2)
    out["total_correlation2_ptw"], out["total_correlation2_avg"] = ptw, avg
    # a sub-selection of channels in non-sorted order
    sel = (5, 1, 6, 2)
    for name in ("total_correlation", "dual_total_correlation", "o_information", "s_information"):
        ptw, avg = getattr(knn, name)(data=data, k=3, idxs=sel)
        out[name + "_sel_ptw"], out[name + "_sel_avg"] = ptw, avg
    subs = {f"small{d}": (d,) for d in range(8)}
    subs.update({f"big{d}": tuple(j for j in range(8) if j != d) for d in range(8)})
    out.update(radius_and_counts(utils, data, tuple(range(8)), subs, 4))
    for k in (1, 4):
        ptw, avg = knn.differential_entropy(data=data, k=k, idxs=(0, 3, 7))
        out[f"entropy_k{k}_ptw"], out[f"entropy_k{k}_avg"] = ptw, avg
    ptw, avg = knn.differential_entropy(data=data, k=3, noise_level=0.05, seed=7)
    out["entropy_noise_ptw"], out["entropy_noise_avg"] = ptw, avg
    np.savez_compressed(os.path.join(HERE, "c4_multivariate.npz"), **out)

    n = 3000
    data, call = workloads.c5_info_rate(n=n, seed=0)
    ptw, avg = knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)
    out = {"n": n, "seed": 0, "k": 5, "ptw": ptw, "avg": avg}
    out.update(radius_and_counts(utils, data, tuple(range(8)),
                                 {"x": (0, 1, 2, 3), "y": (4, 5, 6, 7)}, 5))
    series = workloads.c5_series(n=n, seed=0)
    out["sampen"] = np.array([knn.sample_entropy((c,), series) for c in range(4)])
    out["sampen_m3_r015"] = np.array([knn.sample_entropy((c,), series, m=3, r=0.15) for c in range(4)])
    out["sampen_abs_r"] = knn.sample_entropy((0,), series, m=2, r=0.3, normalize_r=False)
    out["cross_sampen_01"] = knn.cross_sample_entropy((0,), (1,), series)
    out["cross_sampen_23_m3"] = knn.cross_sample_entropy((2,), (3,), series, m=3, r=0.25)
    np.savez_compressed(os.path.join(HERE, "c5_temporal.npz"), **out)

    # ---- Kullback-Leibler divergence between two 3-D clouds of different size
    rng = np.random.default_rng(11)
    post = rng.standard_normal((3, 2000))
    prior = 1.3 * rng.standard_normal((3, 2500)) + 0.2
    out = {"seed": 11}
    for k in (1, 3):
        ptw, avg = knn.kullback_leibler_divergence(post, prior, k=k)
        out[f"k{k}_ptw"], out[f"k{k}_avg"] = ptw, avg
    np.savez_compressed(os.path.join(HERE, "kl_divergence.npz"), **out)

    # ---- edge cases: exact duplicates (eps == 0 -> count -1 -> psi(0) = -inf),
    #      k+1 > n (eps = inf -> count n-1), float32 and int32 input
    out = {}
    rng = np.random.default_rng(5)
    base = rng.standard_normal((2, 40))
    dup = np.concatenate([base, base[:, :10], base[:, :10]], axis=1)  # 10 points x3
    with np.errstate(all="ignore"):
        ptw, avg = knn.mutual_information((0,), (1,), k=2, data=dup)
        out["dup_mi_ptw"], out["dup_mi_avg"] = ptw, avg
        out.update({"dup_" + key: v for key, v in radius_and_counts(
            utils, dup, (0, 1), {"x": (0,), "y": (1,)}, 2).items()})
        tiny = rng.standard_normal((2, 4))
        ptw, avg = knn.mutual_information((0,), (1,), k=5, data=tiny)
        out["tiny_mi_ptw"], out["tiny_mi_avg"] = ptw, avg
        out.update({"tiny_" + key: v for key, v in radius_and_counts(
            utils, tiny, (0, 1), {"x": (0,), "y": (1,)}, 5).items()})
    f32 = rng.standard_normal((3, 500)).astype(np.float32)
    ptw, avg = knn.conditional_mutual_information((0,), (1,), (2,), k=3, data=f32)
    out["f32_cmi_ptw"], out["f32_cmi_avg"] = ptw, avg
    i32 = rng.integers(-50, 50, size=(3, 400)).astype(np.int32)
    with np.errstate(all="ignore"):
        ptw, avg = knn.total_correlation(data=i32, k=3)
    out["i32_tc_ptw"], out["i32_tc_avg"] = ptw, avg
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), **out)

    # ---- digamma at every integer the path can feed it (counts+1 up to a few 1e5),
    #      sampled: all of 0..4096 and a log-spaced tail
    from scipy.special import digamma
    args = np.unique(np.concatenate([np.arange(0, 4097), np.unique(np.geomspace(4097, 2**31 - 1, 4000).astype(np.int64))]))
    with np.errstate(all="ignore"):
        np.savez_compressed(os.path.join(HERE, "digamma.npz"), args=args, psi=digamma(args))

    make_mixed()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
