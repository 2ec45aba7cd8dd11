"""Pins the CPU oracle (oracle/ksg_oracle.py) before anything trusts it:

1. the nine known answers of the reference's tests/test_knn.py:29-75 (JIDT / by hand),
2. the hand-computed radii / counts of tests/test_knn_utils.py:10-92,
3. the Antropy constants of tests/test_knn_temporal.py:55-76,
4. tests/golden/*.npz -- outputs of the unmodified reference run in the build
   container (tests/golden/make_golden.py): radii and counts bit-exact, pointwise
   values bit-exact (same fp64 op order), averages to 1e-13.
"""
import os

import numpy as np
import pytest

from oracle import ksg_oracle as orc
from syntropy_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REL = 1e-6  # the reference's own tolerance (tests/test_knn.py:18)


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_reference_known_answers_tie_fixture():
    data = workloads.tie_fixture()
    assert orc.differential_entropy(data, idxs=(0,), k=1)[1] == pytest.approx(5.870524698081722, REL)
    assert orc.differential_entropy(data, idxs=(0,), k=4)[1] == pytest.approx(4.752310791249405, REL)
    assert orc.mutual_information((0,), (1,), 1, data, algorithm=1)[1] == pytest.approx(5.177377517639623, REL)
    assert orc.mutual_information((0,), (1,), 1, data, algorithm=2)[1] == pytest.approx(2.2173775176396227, REL)
    assert orc.conditional_mutual_information((0,), (1,), (2,), 1, data)[1] == pytest.approx(4.479205338329424, REL)
    assert orc.total_correlation(data, 1)[1] == pytest.approx(5.875549696949821, REL)
    assert orc.dual_total_correlation(data, 1)[1] == pytest.approx(5.1773775176396235, REL)
    assert orc.o_information(data, 1)[1] == pytest.approx(0.6981721793101983, REL)
    assert orc.s_information(data, 1)[1] == pytest.approx(11.052927214589442, REL)


def test_reference_utils_hand_values():
    # tests/test_knn_utils.py:10-30
    eps = orc.knn_radius(np.array([[0.0, 1.0, 2.0, 4.0]]), 1)
    assert same(eps, [1.0, 1.0, 1.0, 2.0])
    # :47-61 Chebyshev distance of (0,0),(3,4) is 4
    assert same(orc.knn_radius(np.array([[0.0, 3.0], [0.0, 4.0]]), 1), [4.0, 4.0])
    # :64-77
    pts = np.array([[0.0, 1.0, 2.0, 5.0]])
    assert same(orc.count_within(pts, np.full(4, 1.5)), [1, 2, 1, 0])
    assert orc.count_within(np.array([[0.0]]), np.array([100.0]))[0] == 0
    # :80-92 strict excludes points exactly on the boundary
    assert same(orc.count_within(pts, np.full(4, 1.0), strict=True), [0, 0, 0, 0])
    assert same(orc.count_within(pts, np.full(4, 1.0), strict=False), [1, 2, 1, 0])


def test_reference_antropy_constants():
    # tests/test_knn_temporal.py:55-76 (first series of each family; the rest are
    # covered through the golden file at smaller n)
    rng = np.random.default_rng(0)
    noise = rng.standard_normal((5, 10_000))
    assert orc.sample_entropy((0,), noise) == pytest.approx(2.1960718154968815, abs=1e-6)
    assert orc.sample_entropy((3,), noise) == pytest.approx(2.1764155747815264, abs=1e-6)
    brownian = np.cumsum(noise, axis=-1)
    assert orc.sample_entropy((0,), brownian) == pytest.approx(0.04959120707043105, abs=1e-6)
    assert orc.sample_entropy((4,), brownian) == pytest.approx(0.0437523057152039, abs=1e-6)


def test_cross_sample_entropy_symmetry_and_brute_force():
    rng = np.random.default_rng(0)
    data = rng.standard_normal((2, 250))
    a = orc.cross_sample_entropy((0,), (1,), data)
    b = orc.cross_sample_entropy((1,), (0,), data)
    assert a == pytest.approx(b, abs=1e-12)


def test_psi_matches_scipy_golden(golden):
    g = golden("digamma")
    with np.errstate(all="ignore"):
        mine = orc.psi_int(g["args"])
    assert same(mine, g["psi"])


def test_golden_tie_fixture(golden):
    g = golden("tie_fixture")
    data = workloads.tie_fixture()
    for k in (1, 4):
        ptw, avg = orc.differential_entropy(data, idxs=(0,), k=k)
        assert same(ptw, g[f"entropy_k{k}_ptw"]) and avg == g[f"entropy_k{k}_avg"]
    ptw, avg = orc.mutual_information((0,), (1,), 1, data, algorithm=1)
    assert same(ptw, g["mi1_ptw"]) and avg == g["mi1_avg"]
    # KSG-2 / TC-2 depend on cKDTree's neighbour choice among exact ties; the
    # average pinned by JIDT is tie-invariant, the pointwise array is not.
    assert orc.mutual_information((0,), (1,), 1, data, algorithm=2)[1] == pytest.approx(float(g["mi2_avg"]), REL)
    ptw, avg = orc.conditional_mutual_information((0,), (1,), (2,), 1, data)
    assert same(ptw, g["cmi_ptw"]) and avg == g["cmi_avg"]
    for name in ("total_correlation", "dual_total_correlation", "o_information", "s_information"):
        ptw, avg = getattr(orc, name)(data, 1)
        assert same(ptw, g[name + "_ptw"]), name
        assert avg == g[name + "_avg"], name
    eps = orc.knn_radius(data, 1)
    assert same(eps, g["rc_eps"])
    for name, idxs in {"x": (0,), "y": (1,), "z": (2,), "xy": (0, 1), "yz": (1, 2)}.items():
        assert same(orc.count_within(data[idxs, :], eps), g["rc_counts_" + name]), name


def test_golden_c1(golden):
    g = golden("c1_mi")
    data, call = workloads.c1_readme_mi(n=int(g["n"]), seed=int(g["seed"]))
    eps = orc.knn_radius(data, 5)
    assert same(eps, g["eps"])
    assert same(orc.count_within(data[(0,), :], eps), g["counts_x"])
    assert same(orc.count_within(data[(1,), :], eps), g["counts_y"])
    ptw, avg = orc.mutual_information((0,), (1,), 5, data)
    assert same(ptw, g["ptw"])
    assert avg == g["avg"]
    assert float(g["avg"]) == pytest.approx(0.780399359365535, rel=1e-13)  # SURVEY.md 8c anchor
    ptw2, avg2 = orc.mutual_information((0,), (1,), 5, data, algorithm=2)
    assert same(ptw2, g["ptw_alg2"]) and avg2 == g["avg_alg2"]


def test_golden_c3(golden):
    g = golden("c3_cmi")
    data, call = workloads.c3_cmi(n=int(g["n"]), seed=0)
    eps = orc.knn_radius(data, 4)
    assert same(eps, g["eps"])
    for name, idxs in {"z": (4, 5), "xz": (0, 1, 4, 5), "yz": (2, 3, 4, 5)}.items():
        assert same(orc.count_within(data[idxs, :], eps), g["counts_" + name])
    ptw, avg = orc.conditional_mutual_information((0, 1), (2, 3), (4, 5), 4, data)
    assert same(ptw, g["ptw"]) and avg == g["avg"]


def test_golden_c4(golden):
    g = golden("c4_multivariate")
    data, _ = workloads.c4_oinfo(n=int(g["n"]), seed=0)
    eps = orc.knn_radius(data, 4)
    assert same(eps, g["eps"])
    for d in range(8):
        assert same(orc.count_within(data[(d,), :], eps), g[f"counts_small{d}"])
        rest = tuple(j for j in range(8) if j != d)
        assert same(orc.count_within(data[rest, :], eps), g[f"counts_big{d}"])
    for name in ("total_correlation", "dual_total_correlation", "o_information", "s_information"):
        ptw, avg = getattr(orc, name)(data, 4)
        assert same(ptw, g[name + "_ptw"]), name
        assert avg == g[name + "_avg"], name
        ptw, avg = getattr(orc, name)(data, 3, idxs=(5, 1, 6, 2))
        assert same(ptw, g[name + "_sel_ptw"]), name
        assert avg == g[name + "_sel_avg"], name
    ptw, avg = orc.total_correlation(data, 4, algorithm=2)
    assert same(ptw, g["total_correlation2_ptw"]) and avg == g["total_correlation2_avg"]
    for k in (1, 4):
        ptw, avg = orc.differential_entropy(data, k, idxs=(0, 3, 7))
        assert same(ptw, g[f"entropy_k{k}_ptw"]) and avg == g[f"entropy_k{k}_avg"]
    ptw, avg = orc.differential_entropy(data, 3, noise_level=0.05, seed=7)
    assert same(ptw, g["entropy_noise_ptw"]) and avg == g["entropy_noise_avg"]


def test_golden_c5(golden):
    g = golden("c5_temporal")
    n = int(g["n"])
    data, call = workloads.c5_info_rate(n=n, seed=0)
    eps = orc.knn_radius(data, 5)
    assert same(eps, g["eps"])
    assert same(orc.count_within(data[(0, 1, 2, 3), :], eps), g["counts_x"])
    assert same(orc.count_within(data[(4, 5, 6, 7), :], eps), g["counts_y"])
    ptw, avg = orc.mutual_information((0, 1, 2, 3), (4, 5, 6, 7), 5, data)
    assert same(ptw, g["ptw"]) and avg == g["avg"]
    series = workloads.c5_series(n=n, seed=0)
    assert same([orc.sample_entropy((c,), series) for c in range(4)], g["sampen"])
    assert same([orc.sample_entropy((c,), series, m=3, r=0.15) for c in range(4)], g["sampen_m3_r015"])
    assert orc.sample_entropy((0,), series, m=2, r=0.3, normalize_r=False) == g["sampen_abs_r"]
    assert orc.cross_sample_entropy((0,), (1,), series) == g["cross_sampen_01"]
    assert orc.cross_sample_entropy((2,), (3,), series, m=3, r=0.25) == g["cross_sampen_23_m3"]


def test_golden_kl(golden):
    g = golden("kl_divergence")
    rng = np.random.default_rng(int(g["seed"]))
    post = rng.standard_normal((3, 2000))
    prior = 1.3 * rng.standard_normal((3, 2500)) + 0.2
    for k in (1, 3):
        ptw, avg = orc.kullback_leibler_divergence(post, prior, k)
        assert same(ptw, g[f"k{k}_ptw"]) and avg == g[f"k{k}_avg"]


def test_golden_edge_cases(golden):
    g = golden("edge_cases")
    rng = np.random.default_rng(5)
    base = rng.standard_normal((2, 40))
    dup = np.concatenate([base, base[:, :10], base[:, :10]], axis=1)
    with np.errstate(all="ignore"):
        eps = orc.knn_radius(dup, 2)
        assert same(eps, g["dup_eps"]) and (eps == 0).sum() == 30
        assert same(orc.count_within(dup[(0,), :], eps), g["dup_counts_x"])
        assert orc.count_within(dup[(0,), :], eps).min() == -1
        ptw, avg = orc.mutual_information((0,), (1,), 2, dup)
        assert same(ptw, g["dup_mi_ptw"]) and same(avg, g["dup_mi_avg"])
        tiny = rng.standard_normal((2, 4))
        eps = orc.knn_radius(tiny, 5)
        assert same(eps, g["tiny_eps"]) and np.isinf(eps).all()
        assert same(orc.count_within(tiny[(1,), :], eps), g["tiny_counts_y"])
        ptw, avg = orc.mutual_information((0,), (1,), 5, tiny)
        assert same(ptw, g["tiny_mi_ptw"]) and same(avg, g["tiny_mi_avg"])
    f32 = rng.standard_normal((3, 500)).astype(np.float32)
    ptw, avg = orc.conditional_mutual_information((0,), (1,), (2,), 3, f32)
    assert same(ptw, g["f32_cmi_ptw"]) and avg == g["f32_cmi_avg"]
    i32 = rng.integers(-50, 50, size=(3, 400)).astype(np.int32)
    with np.errstate(all="ignore"):
        ptw, avg = orc.total_correlation(i32, 3)
    assert same(ptw, g["i32_tc_ptw"]) and same(avg, g["i32_tc_avg"])


def test_c_restatement_matches_numpy_oracle():
    """oracle/ksg_oracle.c (gcc, OpenMP) == oracle/ksg_oracle.py on radii, counts and pair counts."""
    from oracle import ksg_oracle_c as oc

    rng = np.random.default_rng(8)
    rows = rng.standard_normal((5, 1500))
    eps = orc.knn_radius(rows, 4)
    assert same(oc.knn_radius(rows, 5), eps)
    assert same(oc.knn_radius(rows, 3, queries=rows[:, :100] + 0.01), orc.knn_radius(rows, 3, queries=rows[:, :100] + 0.01))
    for sub in ((0,), (1, 3), (0, 1, 2, 3, 4)):
        assert same(oc.count_within(rows[sub, :], eps), orc.count_within(rows[sub, :], eps))
        assert same(oc.count_within(rows[sub, :], eps, strict=False), orc.count_within(rows[sub, :], eps, strict=False))
    x, y = rng.standard_normal(800), rng.standard_normal(800)
    for m in (1, 2, 3):
        nt = x.size - m
        small = orc.pair_count(orc._embed(x, m, nt), orc._embed(y, m, nt), 0.4)
        big = orc.pair_count(orc._embed(x, m + 1, nt), orc._embed(y, m + 1, nt), 0.4)
        assert oc.pair_counts(x, y, m, 0.4) == (small, big)


def test_mixed_knn_branch_vs_reference_golden(golden):
    """syntropy.mixed with continuous_estimator="knn" (SURVEY.md 8f rank 3): the restatement
    reproduces the unmodified reference on every fixture of workloads.mixed_cases() -- pointwise
    arrays to the last bit (same float64 operation order), averages to 1e-13."""
    g = golden("mixed_knn")
    with np.errstate(all="ignore"):
        for name, (d, c, k) in workloads.mixed_cases().items():
            ptw, avg = orc.mixed_shannon_entropy(d, c, k)
            assert same(ptw, g[f"{name}_h_ptw"]) and same(avg, g[f"{name}_h_avg"])
            for cond in ("discrete", "continuous"):
                ptw, avg = orc.mixed_conditional_entropy(d, c, cond, k)
                assert same(ptw, g[f"{name}_ce_{cond}_ptw"])
                assert avg == pytest.approx(g[f"{name}_ce_{cond}_avg"], rel=1e-13) or same(avg, g[f"{name}_ce_{cond}_avg"])
            ptw, avg = orc.mixed_mutual_information(d, c, k)
            assert same(ptw, g[f"{name}_mi_ptw"])
            assert avg == pytest.approx(g[f"{name}_mi_avg"], rel=1e-13) or same(avg, g[f"{name}_mi_avg"])


# ------------------------------------------------------------------ Minkowski p
def _minkowski_inputs():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.minkowski_cases()


def test_oracle_minkowski_p_vs_reference_golden(golden):
    """p = 1 and p = 2 (cKDTree's p-th-power-space arithmetic incl. the four-way unrolled sum of
    squares) and p = 3 (libm pow through the C restatement) reproduce the unmodified reference bit
    for bit: radii, strict / closed counts, every estimator that takes ``p`` / ``norm``, tied data."""
    g = golden("minkowski")
    cont, tied, post, prior, series = _minkowski_inputs()
    with np.errstate(all="ignore"):
        for p in (1, 2, 3):
            t = f"p{p}"
            eps = orc.knn_radius_p(cont[(0, 1, 2), :], 4, p)
            assert np.array_equal(eps, g[f"{t}_eps"])
            assert np.array_equal(orc.count_within_p(cont[(0, 1), :], eps, p), g[f"{t}_counts_strict"])
            assert np.array_equal(orc.count_within_p(cont[(0, 1), :], eps, p, strict=False), g[f"{t}_counts_closed"])
            ptw, avg = orc.differential_entropy(cont, 4, idxs=(0, 2, 5), p=p)
            assert np.array_equal(ptw, g[f"{t}_entropy_ptw"]) and avg == g[f"{t}_entropy_avg"]
            for alg in (1, 2):
                ptw, avg = orc.mutual_information((0, 1), (2, 3), 4, cont, algorithm=alg, p=p)
                assert np.array_equal(ptw, g[f"{t}_mi{alg}_ptw"]), (p, alg)
                assert avg == g[f"{t}_mi{alg}_avg"]
            ptw, avg = orc.mutual_information((0,), (1,), 3, cont, algorithm=1, p=p)
            assert np.array_equal(ptw, g[f"{t}_mi1_1d_ptw"])
            ptw, avg = orc.conditional_mutual_information((0, 1), (2,), (3, 4), 3, cont, p=p)
            assert np.array_equal(ptw, g[f"{t}_cmi_ptw"]) and avg == g[f"{t}_cmi_avg"]
            ptw, avg = orc.kullback_leibler_divergence(post, prior, 3, p=p)
            assert np.array_equal(ptw, g[f"{t}_kl_ptw"]) and avg == g[f"{t}_kl_avg"]
            se = np.array([orc.sample_entropy((c,), series, m=2, r=0.3, norm=p) for c in range(2)])
            assert np.array_equal(se, g[f"{t}_sampen"])
            assert orc.cross_sample_entropy((0,), (1,), series, m=2, r=0.35, norm=p) == g[f"{t}_cross_sampen"]
            if p != 3:
                ptw, avg = orc.mutual_information((0, 1), (2,), 3, tied, algorithm=1, p=p)
                assert np.array_equal(ptw, g[f"{t}_tied_mi1_ptw"], equal_nan=True)
                ptw, avg = orc.conditional_mutual_information((0,), (1,), (2, 3), 2, tied, p=p)
                assert np.array_equal(ptw, g[f"{t}_tied_cmi_ptw"], equal_nan=True)   # (psi(0) - psi(0) = nan)
