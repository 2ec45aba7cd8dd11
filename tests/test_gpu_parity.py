"""GPU parity tests proper (run with ``-m gpu`` on the B200 box).

Everything goes through the C ABI of include/ksg_b200.h (via syntropy_b200._engine /
the public syntropy_b200.knn functions) and is compared with
 (a) the committed golden fixtures produced by the unmodified reference, and
 (b) the CPU oracle on the same seeded inputs.
Bars: radii and neighbour counts bit-exact; pointwise / average nats within 1e-12
relative (the only non-bit-exact ingredient is log() inside digamma for arguments
above 10, where CUDA's libdevice and the host libm may differ in the last ulp).
"""
import os

import numpy as np
import pytest

from oracle import ksg_oracle as orc
from syntropy_b200 import workloads

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(scope="module")
def knn(E):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from syntropy_b200 import knn as mod

    return mod


@pytest.fixture(scope="module", params=["windowed", "dense", "exact"])
def E(request):
    """Every test runs under each engine: the fp32-filtered kernels with and without
    window pruning, and the float64 dense kernels.  All three must give the same bits."""
    from syntropy_b200 import _engine

    old = _engine.config.engine
    _engine.config.engine = request.param
    yield _engine
    _engine.config.engine = old


def ptw_floor(ref_ptw):
    """mean |ptw_ref| over the finite entries: the floor SURVEY.md 8(d) allows under the 1e-12
    relative bar (pointwise values are sums of psi terms of this size that may cancel to ~0)."""
    b = np.asarray(ref_ptw, dtype=np.float64).reshape(-1)
    b = b[np.isfinite(b)]
    return float(np.mean(np.abs(b))) if b.size else 0.0


def close(a, b, floor=None):
    """|a-b| <= 1e-12 * max(|b|, floor).  ``floor`` defaults to mean |b| for an array and to 0
    (purely relative) for a scalar; an average is compared with the floor of its own pointwise
    reference (``close_pair``)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if floor is None:
        floor = ptw_floor(b) if b.size > 1 else 0.0
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    ok = both_inf | both_nan | (np.abs(a - b) <= RTOL * np.maximum(np.abs(b), floor))
    assert ok.all(), f"max abs diff {np.nanmax(np.abs(a - b)[~(both_inf | both_nan)])}, floor {floor}"


def close_pair(ptw, ref_ptw, avg, ref_avg):
    """Pointwise array and average of one estimator call against the reference's."""
    close(ptw, ref_ptw)
    close(avg, ref_avg, floor=ptw_floor(ref_ptw))


def radius_and_counts(E, rows, k, subspaces, strict=True, fused=False):
    import torch

    pts = E.to_device_points(rows)
    dim, n = pts.shape
    masks = [E.mask_of(s) for s in subspaces]
    if not E._use_fast(dim, k + 1):
        eps = E.knn_radius(pts, k + 1, 0, n)
        counts = E.count_subspaces(pts, masks, eps, 0, n, strict=strict)
    elif strict and not fused:
        # the product path: joint-order passes + low-dimensional marginals in their own order
        eps, counts, _ = E.radius_and_counts(pts, k, masks)
    else:
        windowed = E.config.engine == "windowed"
        prep = E.PreparedSet(pts, E.config.primary)
        eps_s = E.fast_knn(prep, k + 1, 0, n, windowed)
        counts_s = E.fast_counts(prep, masks, eps_s, 0, n, windowed, strict=strict)
        perm = prep.perm.long()
        eps = torch.empty_like(eps_s)
        eps[perm] = eps_s
        counts = torch.empty_like(counts_s)
        counts[:, perm] = counts_s
    torch.cuda.synchronize()
    return eps.cpu().numpy(), counts.cpu().numpy().astype(np.int64)


# ------------------------------------------------------------------ golden fixtures
def test_psi_table_vs_scipy_golden(E, golden):
    g = golden("digamma")
    args = g["args"]
    small = args[args < 3_000_000]
    table = E.psi_table(int(small.max()) + 1).cpu().numpy()
    ref = g["psi"][: small.size]
    mine = table[small]
    assert np.isneginf(mine[0])
    assert np.array_equal(mine[small <= 10], ref[small <= 10])  # harmonic sums: bit-exact
    rel = np.abs(mine[1:] - ref[1:]) / np.abs(ref[1:])
    assert rel.max() < 4e-16


def test_golden_tie_fixture(knn, E, golden):
    g = golden("tie_fixture")
    data = workloads.tie_fixture()
    for k in (1, 4):
        ptw, avg = knn.differential_entropy(data=data, idxs=(0,), k=k)
        close_pair(ptw, g[f"entropy_k{k}_ptw"], avg, g[f"entropy_k{k}_avg"])
    ptw, avg = knn.mutual_information((0,), (1,), k=1, data=data, algorithm=1)
    close_pair(ptw, g["mi1_ptw"], avg, g["mi1_avg"])
    assert avg == pytest.approx(5.177377517639623, 1e-6)   # tests/test_knn.py:44 (JIDT)
    mi2 = knn.mutual_information((0,), (1,), k=1, data=data, algorithm=2)[1]
    assert mi2 == pytest.approx(2.2173775176396227, 1e-6)  # tests/test_knn.py:50 (JIDT)
    ptw, avg = knn.conditional_mutual_information((0,), (1,), (2,), k=1, data=data)
    close_pair(ptw, g["cmi_ptw"], avg, g["cmi_avg"])
    assert avg == pytest.approx(4.479205338329424, 1e-6)   # tests/test_knn.py:57
    known = {"total_correlation": 5.875549696949821, "dual_total_correlation": 5.1773775176396235,
             "o_information": 0.6981721793101983, "s_information": 11.052927214589442}
    for name, val in known.items():
        ptw, avg = getattr(knn, name)(data=data, k=1)
        close_pair(ptw, g[name + "_ptw"], avg, g[name + "_avg"])
        assert avg == pytest.approx(val, 1e-6)               # tests/test_knn.py:60-75
    subs = {"x": (0,), "y": (1,), "z": (2,), "xy": (0, 1), "yz": (1, 2)}
    eps, counts = radius_and_counts(E, data, 1, list(subs.values()))
    assert np.array_equal(eps, g["rc_eps"])
    for row, name in enumerate(subs):
        assert np.array_equal(counts[row], g["rc_counts_" + name]), name


def test_golden_c1_readme_case(knn, E, golden):
    g = golden("c1_mi")
    data, call = workloads.c1_readme_mi(n=int(g["n"]), seed=0)
    eps, counts = radius_and_counts(E, data, 5, [(0,), (1,)])
    assert np.array_equal(eps, g["eps"])
    assert np.array_equal(counts[0], g["counts_x"])
    assert np.array_equal(counts[1], g["counts_y"])
    ptw, avg = knn.mutual_information((0,), (1,), k=5, data=data)
    assert ptw.shape == (1, 5000) and isinstance(avg, np.float64)
    close_pair(ptw, g["ptw"], avg, g["avg"])
    assert avg == pytest.approx(0.780399359365535, rel=1e-12)
    ptw, avg = knn.mutual_information((0,), (1,), k=5, data=data, algorithm=2)
    close_pair(ptw, g["ptw_alg2"], avg, g["avg_alg2"])


def test_golden_c3_cmi(knn, E, golden):
    g = golden("c3_cmi")
    data, call = workloads.c3_cmi(n=int(g["n"]), seed=0)
    eps, counts = radius_and_counts(E, data, 4, [(4, 5), (0, 1, 4, 5), (2, 3, 4, 5)])
    assert np.array_equal(eps, g["eps"])
    for row, name in enumerate(("z", "xz", "yz")):
        assert np.array_equal(counts[row], g["counts_" + name]), name
    ptw, avg = knn.conditional_mutual_information((0, 1), (2, 3), (4, 5), k=4, data=data)
    close_pair(ptw, g["ptw"], avg, g["avg"])


def test_golden_c4_multivariate(knn, E, golden):
    g = golden("c4_multivariate")
    data, _ = workloads.c4_oinfo(n=int(g["n"]), seed=0)
    subs = [(d,) for d in range(8)] + [tuple(j for j in range(8) if j != d) for d in range(8)]
    eps, counts = radius_and_counts(E, data, 4, subs)
    assert np.array_equal(eps, g["eps"])
    for d in range(8):
        assert np.array_equal(counts[d], g[f"counts_small{d}"])
        assert np.array_equal(counts[8 + d], g[f"counts_big{d}"])
    for name in ("total_correlation", "dual_total_correlation", "o_information", "s_information"):
        ptw, avg = getattr(knn, name)(data=data, k=4)
        close_pair(ptw, g[name + "_ptw"], avg, g[name + "_avg"])
        ptw, avg = getattr(knn, name)(data=data, k=3, idxs=(5, 1, 6, 2))
        close_pair(ptw, g[name + "_sel_ptw"], avg, g[name + "_sel_avg"])
    ptw, avg = knn.total_correlation(data=data, k=4, algorithm=2)
    close_pair(ptw, g["total_correlation2_ptw"], avg, g["total_correlation2_avg"])
    for k in (1, 4):
        ptw, avg = knn.differential_entropy(data=data, k=k, idxs=(0, 3, 7))
        close_pair(ptw, g[f"entropy_k{k}_ptw"], avg, g[f"entropy_k{k}_avg"])
    ptw, avg = knn.differential_entropy(data=data, k=3, noise_level=0.05, seed=7)
    close_pair(ptw, g["entropy_noise_ptw"], avg, g["entropy_noise_avg"])
    # identities O = TC - DTC, S = TC + DTC hold on the outputs
    tc = knn.total_correlation(data, 4)[1]; dtc = knn.dual_total_correlation(data, 4)[1]
    assert knn.o_information(data, 4)[1] == pytest.approx(tc - dtc, abs=1e-9)
    assert knn.s_information(data, 4)[1] == pytest.approx(tc + dtc, abs=1e-9)


def test_golden_c5_and_temporal(knn, E, golden):
    g = golden("c5_temporal")
    n = int(g["n"])
    data, call = workloads.c5_info_rate(n=n, seed=0)
    eps, counts = radius_and_counts(E, data, 5, [(0, 1, 2, 3), (4, 5, 6, 7)])
    assert np.array_equal(eps, g["eps"])
    assert np.array_equal(counts[0], g["counts_x"]) and np.array_equal(counts[1], g["counts_y"])
    ptw, avg = knn.mutual_information((0, 1, 2, 3), (4, 5, 6, 7), k=5, data=data)
    close_pair(ptw, g["ptw"], avg, g["avg"])
    series = workloads.c5_series(n=n, seed=0)
    # pair counts are integers: the entropies are bit-exact up to the host log
    assert [knn.sample_entropy((c,), series) for c in range(4)] == list(g["sampen"])
    assert [knn.sample_entropy((c,), series, m=3, r=0.15) for c in range(4)] == list(g["sampen_m3_r015"])
    assert knn.sample_entropy((0,), series, m=2, r=0.3, normalize_r=False) == g["sampen_abs_r"]
    assert knn.cross_sample_entropy((0,), (1,), series) == g["cross_sampen_01"]
    assert knn.cross_sample_entropy((2,), (3,), series, m=3, r=0.25) == g["cross_sampen_23_m3"]


def test_golden_kl_divergence(knn, golden):
    g = golden("kl_divergence")
    rng = np.random.default_rng(int(g["seed"]))
    post = rng.standard_normal((3, 2000))
    prior = 1.3 * rng.standard_normal((3, 2500)) + 0.2
    for k in (1, 3):
        ptw, avg = knn.kullback_leibler_divergence(post, prior, k=k)
        assert ptw.shape == (2000,)
        close_pair(ptw, g[f"k{k}_ptw"], avg, g[f"k{k}_avg"])


def test_golden_edge_cases(knn, E, golden):
    g = golden("edge_cases")
    rng = np.random.default_rng(5)
    base = rng.standard_normal((2, 40))
    dup = np.concatenate([base, base[:, :10], base[:, :10]], axis=1)
    eps, counts = radius_and_counts(E, dup, 2, [(0,), (1,)])
    assert np.array_equal(eps, g["dup_eps"]) and (eps == 0).sum() == 30
    assert np.array_equal(counts[0], g["dup_counts_x"]) and counts[0].min() == -1
    assert np.array_equal(counts[1], g["dup_counts_y"])
    ptw, avg = knn.mutual_information((0,), (1,), k=2, data=dup)
    close_pair(ptw, g["dup_mi_ptw"], avg, g["dup_mi_avg"])     # +inf entries from psi(0)
    tiny = rng.standard_normal((2, 4))
    eps, counts = radius_and_counts(E, tiny, 5, [(0,), (1,)])     # k+1 > n -> eps = inf
    assert np.isinf(eps).all() and np.array_equal(counts[0], g["tiny_counts_x"])
    ptw, avg = knn.mutual_information((0,), (1,), k=5, data=tiny)
    close_pair(ptw, g["tiny_mi_ptw"], avg, g["tiny_mi_avg"])
    f32 = rng.standard_normal((3, 500)).astype(np.float32)
    ptw, avg = knn.conditional_mutual_information((0,), (1,), (2,), k=3, data=f32)
    close_pair(ptw, g["f32_cmi_ptw"], avg, g["f32_cmi_avg"])
    i32 = rng.integers(-50, 50, size=(3, 400)).astype(np.int32)
    ptw, avg = knn.total_correlation(data=i32, k=3)
    close_pair(ptw, g["i32_tc_ptw"], avg, g["i32_tc_avg"])


# ------------------------------------------------------------------ reference-test semantics
def test_reference_utils_semantics(E):
    # tests/test_knn_utils.py:10-30, 64-92 restated at the C-ABI level
    eps, _ = radius_and_counts(E, np.array([[0.0, 1.0, 2.0, 4.0]]), 1, [(0,)])
    assert np.array_equal(eps, [1.0, 1.0, 1.0, 2.0])
    import torch
    pts = E.to_device_points(np.array([[0.0, 1.0, 2.0, 5.0]]))
    for radius, strict, want in ((1.5, True, [1, 2, 1, 0]), (1.0, True, [0, 0, 0, 0]), (1.0, False, [1, 2, 1, 0])):
        e = torch.full((4,), radius, dtype=torch.float64, device=pts.device)
        c = E.count_subspaces(pts, [1], e, 0, 4, strict=strict).cpu().numpy()[0]
        assert np.array_equal(c, want), (radius, strict)
    single = E.to_device_points(np.array([[0.0]]))
    e = torch.full((1,), 100.0, dtype=torch.float64, device=pts.device)
    assert E.count_subspaces(single, [1], e, 0, 1).cpu().numpy()[0, 0] == 0


def test_errors_mirror_reference(knn):
    data = np.random.default_rng(0).standard_normal((3, 64))
    with pytest.raises(AssertionError):
        knn.mutual_information((0,), (1,), 3, data, algorithm=3)
    with pytest.raises(AssertionError):
        knn.total_correlation(data, 3, algorithm=0)
    with pytest.raises(AssertionError):
        knn.kullback_leibler_divergence(data, data[:2], 3)
    bad = data.copy(); bad[1, 5] = np.nan
    with pytest.raises(ValueError):
        knn.mutual_information((0,), (1,), 3, bad)
    with pytest.raises(ValueError):
        knn.mutual_information((0,), (1,), 3, data, p=0.5)  # cKDTree: only p-norms with 1 <= p <= inf
    ptw, avg = knn.o_information(data[:2], 3)
    assert ptw.shape == (64,) and avg == 0.0 and not ptw.any()


# ------------------------------------------------------------------ oracle at moderate n
@pytest.mark.parametrize("dim,k,n", [(1, 1, 777), (2, 5, 6000), (3, 3, 4097), (5, 4, 3000), (8, 4, 5000), (11, 2, 1500),
                                     (2, 12, 5000), (3, 20, 4000), (6, 30, 2500), (2, 40, 3000)])
def test_radius_counts_vs_oracle(E, dim, k, n):
    rng = np.random.default_rng(100 + dim)
    rows = rng.standard_normal((dim, n)) * rng.uniform(0.5, 3.0, size=(dim, 1))
    subs = [(d,) for d in range(dim)]
    if dim > 1:
        subs += [tuple(range(dim // 2)), tuple(range(dim // 2, dim)), tuple(range(1, dim))]
    eps, counts = radius_and_counts(E, rows, k, subs)
    ref_eps = orc.knn_radius(rows, k)
    assert np.array_equal(eps, ref_eps)
    for row, s in enumerate(subs):
        assert np.array_equal(counts[row], orc.count_within(rows[list(s), :], ref_eps)), s
    _, closed = radius_and_counts(E, rows, k, subs[:2], strict=False)
    assert np.array_equal(closed[0], orc.count_within(rows[list(subs[0]), :], ref_eps, strict=False))


def test_quantised_data_ties(E, knn):
    rng = np.random.default_rng(3)
    rows = np.round(rng.standard_normal((3, 3000)) * 4) / 4  # heavy exact ties
    eps, counts = radius_and_counts(E, rows, 3, [(0,), (1, 2), (0, 1, 2)])
    ref = orc.knn_radius(rows, 3)
    assert np.array_equal(eps, ref)
    assert np.array_equal(counts[0], orc.count_within(rows[(0,), :], ref))
    assert np.array_equal(counts[1], orc.count_within(rows[(1, 2), :], ref))
    assert np.array_equal(counts[2], orc.count_within(rows, ref))
    with np.errstate(all="ignore"):
        close(knn.total_correlation(rows, 3)[0], orc.total_correlation(rows, 3)[0])


def test_query_slices_compose(E):
    """Sharding invariant: any query slice equals the same rows of the full result."""
    rows = np.random.default_rng(9).standard_normal((4, 5000))
    pts = E.to_device_points(rows)
    full = E.knn_radius(pts, 5, 0, 5000)
    part = E.knn_radius(pts, 5, 1234, 3999)
    assert np.array_equal(full[1234:3999].cpu().numpy(), part.cpu().numpy())
    cf = E.count_subspaces(pts, [3, 12], full, 0, 5000)
    cp = E.count_subspaces(pts, [3, 12], part, 1234, 3999)
    assert np.array_equal(cf[:, 1234:3999].cpu().numpy(), cp.cpu().numpy())


def test_knn_indices_vs_oracle(E):
    rows = np.random.default_rng(21).standard_normal((3, 2500))
    pts = E.to_device_points(rows)
    eps, idx = E.knn_radius(pts, 5, 0, 2500, want_indices=True)
    ref_eps, ref_idx = orc.knn_radius(rows, 4, return_indices=True)
    assert np.array_equal(eps.cpu().numpy(), ref_eps)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)


def test_pair_counts_vs_oracle(E):
    rng = np.random.default_rng(4)
    x = rng.standard_normal(3000); y = rng.standard_normal(3000)
    for m in (1, 2, 3, 7):
        nt = x.size - m
        small, big = E.pair_counts(x, y, m, 0.35)
        assert small == orc.pair_count(orc._embed(x, m, nt), orc._embed(y, m, nt), 0.35)
        assert big == orc.pair_count(orc._embed(x, m + 1, nt), orc._embed(y, m + 1, nt), 0.35)


def test_deterministic_sum(E):
    import torch
    v = np.random.default_rng(0).standard_normal(1_000_003)
    t = torch.from_numpy(v).cuda()
    s1 = E.device_sum(t).item(); s2 = E.device_sum(t).item()
    assert s1 == s2
    assert s1 == pytest.approx(float(np.sum(v)), rel=1e-13)


# ------------------------------------------------------------------ full-size properties (C2)
def test_c2_full_size(E, knn):
    """n = 1e6 (the metric's configuration): oracle on a random query subset against the
    full point set, plus size-independent properties of the counts."""
    data, call = workloads.c2_mi_1m()
    n = data.shape[1]
    eps, counts = radius_and_counts(E, data, 5, [(0,), (1,), (0, 1)])
    sub = np.sort(np.random.default_rng(1).choice(n, size=768, replace=False))
    ref = orc.knn_radius(data, 6, queries=data[:, sub])  # k' = 6 = k+1, self included
    assert np.array_equal(eps[sub], ref)
    for row, s in enumerate(((0,), (1,))):
        want = orc.count_within(data[list(s), :], ref, queries=data[list(s), :][:, sub])
        assert np.array_equal(counts[row, sub], want)
    # property: in the joint space exactly k-1 points other than the query lie strictly
    # inside the (k+1)-th-neighbour radius unless distances tie (continuous data: never)
    assert (counts[2] == 4).all()
    # property: marginal counts dominate the joint count
    assert (counts[0] >= counts[2]).all() and (counts[1] >= counts[2]).all()
    ptw, avg = knn.mutual_information((0,), (1,), 5, data)
    assert avg == pytest.approx(0.8044789267532502, rel=1e-12)  # SURVEY.md 8c anchor (reference, n=1e6)
    check_full_size(ptw, avg, _full_size_golden()["c2"]["mi"])


def test_engines_agree_bitwise_on_random_configs():
    """exact fp64, dense fp32-filter and windowed fp32-filter (curve order, kd order and
    single-coordinate order of the prepared set) give identical radii / counts."""
    from syntropy_b200 import _engine as E

    rng = np.random.default_rng(77)
    old, old_primary = E.config.engine, E.config.primary
    try:
        for dim, n, k in ((2, 30_000, 5), (4, 20_000, 3), (6, 12_000, 4), (7, 9_000, 2), (8, 10_000, 4)):
            rows = rng.standard_normal((dim, n)) + 0.3 * rng.standard_normal((1, n))
            subs = [(d,) for d in range(dim)] + [tuple(j for j in range(dim) if j != d) for d in range(dim)]
            if dim >= 4:
                subs += [tuple(range(dim // 2)), tuple(range(dim // 2, dim)), (0, dim - 1)]
            got = {}
            for engine, primary in (("exact", -1), ("dense", -1), ("windowed", -1), ("windowed", -2), ("dense", -2),
                                    ("windowed", dim - 1)):
                E.config.engine, E.config.primary = engine, primary
                got[(engine, primary)] = radius_and_counts(E, rows, k, subs)
            for key in got:
                assert np.array_equal(got[key][0], got[("exact", -1)][0]), (dim, key, "eps")
                assert np.array_equal(got[key][1], got[("exact", -1)][1]), (dim, key, "counts")
    finally:
        E.config.engine, E.config.primary = old, old_primary


def test_own_order_counts_equal_fused_joint_order_counts(E):
    """Every multi-dimensional marginal counted in a prepared set of its own coordinates gives
    the counts of the fused joint-order pass (and of the exact engine, tested above)."""
    for name, n in (("c3", 6000), ("c5", 5000)):
        data, call = workloads.CONFIGS[name](n=n, seed=5)
        if name == "c3":
            subs = [(4, 5), (0, 1, 4, 5), (2, 3, 4, 5)]
        else:
            subs = [(0, 1, 2, 3), (4, 5, 6, 7)]
        eps_a, counts_a = radius_and_counts(E, data, call["k"], subs)
        eps_b, counts_b = radius_and_counts(E, data, call["k"], subs, fused=True)
        assert np.array_equal(eps_a, eps_b)
        assert np.array_equal(counts_a, counts_b)


def test_windowed_pair_counts_equal_dense_fp64_pair_counts(E):
    """sample-entropy pair counts from prepared template sets (self counts; cross counts as
    (UU - XX - YY) / 2) equal the dense float64 kernel's, also when many distances hit the
    radius exactly (quantised series, closed ball)."""
    import torch

    rng = np.random.default_rng(12)
    n = 40_000
    x = np.cumsum(rng.standard_normal(n)) * 0.05 + rng.standard_normal(n)
    y = 0.5 * x + rng.standard_normal(n)
    xq = np.round(x * 4) / 4  # multiples of 0.25: distances equal to r = 0.5 are common
    for a, b, m, r in ((x, x, 2, 0.2 * x.std()), (x, y, 2, 0.3), (xq, xq, 2, 0.5), (xq, np.round(y * 4) / 4, 3, 0.5)):
        got = E.pair_counts(a, a if b is a else b, m, r)
        xs = torch.from_numpy(a).cuda()
        ys = xs if b is a else torch.from_numpy(b).cuda()
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        want = E._pair_counts_dense(xs, ys, n - m, m, r, flag)
        assert got == want


@pytest.mark.parametrize("k", [15, 45, 70])
def test_large_k_through_the_api(knn, k):
    """k beyond what the shared-memory neighbour lists hold falls back to the float64 kernels
    (k + 1 > 64 included): same numbers as the oracle either way."""
    data, _ = workloads.c3_cmi(n=1500, seed=9)
    ptw, avg = knn.mutual_information((0, 1), (2,), k, data)
    ref_ptw, ref_avg = orc.mutual_information((0, 1), (2,), k, data)
    close_pair(ptw, ref_ptw, avg, ref_avg)


def test_k_beyond_the_kernel_limit_is_refused_loudly(knn):
    data, _ = workloads.c1_readme_mi(n=600, seed=2)
    with pytest.raises(NotImplementedError):
        knn.mutual_information((0,), (1,), 200, data)


def test_ksg2_vs_oracle_multidimensional_marginals(knn):
    """KSG algorithm 2 through the fast path (neighbour identities from the fp32 filter + exact
    refine, per-subspace radii, closed-ball counts in each marginal's own prepared set) equals the
    oracle on tie-free data: MI 2D;2D, MI 1D;3D, TC-2 over 5 variables."""
    data, _ = workloads.c3_cmi(n=5000, seed=11)
    for ix, iy, k in (((0, 1), (2, 3), 4), ((4,), (0, 2, 5), 3)):
        ptw, avg = knn.mutual_information(ix, iy, k, data, algorithm=2)
        ref_ptw, ref_avg = orc.mutual_information(ix, iy, k, data, algorithm=2)
        close_pair(ptw, ref_ptw, avg, ref_avg)
    ptw, avg = knn.total_correlation(data[:5], 4, algorithm=2)
    ref_ptw, ref_avg = orc.total_correlation(data[:5], 4, algorithm=2)
    close_pair(ptw, ref_ptw, avg, ref_avg)


@pytest.mark.parametrize("dim,n,m,k,shift", [(3, 20_000, 15_000, 1, 0.5), (2, 9_000, 30_000, 4, 0.0),
                                             (5, 6_000, 6_000, 3, 2.0), (1, 5_000, 7_000, 2, 0.3),
                                             (3, 4_000, 3_000, 1, 25.0)])
def test_kl_divergence_cross_set_search_vs_oracle(knn, dim, n, m, k, shift):
    """The prior-sample radii come from a cross-set search (posterior points placed on the prior's
    curve); different sample sizes, overlapping and far-apart clouds (shift = 25: every query lies
    outside the prior's bounding box)."""
    rng = np.random.default_rng(1000 + dim)
    post = rng.standard_normal((dim, n)) * rng.uniform(0.5, 2.0, size=(dim, 1)) + shift
    prior = rng.standard_normal((dim, m)) * 1.3
    ptw, avg = knn.kullback_leibler_divergence(post, prior, k)
    ref_ptw, ref_avg = orc.kullback_leibler_divergence(post, prior, k)
    close_pair(ptw, ref_ptw, avg, ref_avg)


# ------------------------------------------------------------------ full-size properties (C3, C4, C5)
@pytest.fixture()
def windowed():
    """(engine module, knn module) with the production engine selected: the full BASELINE sizes are
    run on it only (the dense engines need minutes there; they are compared at moderate n above)."""
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from syntropy_b200 import _engine, knn as mod

    old = _engine.config.engine
    _engine.config.engine = "windowed"
    yield _engine, mod
    _engine.config.engine = old


def _subset_parity(E, rows, k, subs, n_check=256, seed=1):
    """radii / counts of ALL samples from the product path; oracle (C restatement, brute force in
    float64 against the full point set) on a random query subset."""
    from oracle import ksg_oracle_c as oc

    n = rows.shape[1]
    eps, counts = radius_and_counts(E, rows, k, subs)
    sel = np.sort(np.random.default_rng(seed).choice(n, size=n_check, replace=False))
    ref = oc.knn_radius(rows, k + 1, queries=rows[:, sel])
    assert np.array_equal(eps[sel], ref)
    for row, s in enumerate(subs):
        want = oc.count_within(rows[list(s), :], ref, queries=rows[list(s), :][:, sel])
        assert np.array_equal(counts[row, sel], want), s
    return eps, counts


def _full_size_golden():
    import json

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_size.json")))


def check_full_size(ptw, avg, g):
    """Average and 64 probed pointwise values against tests/golden/full_size.json (fast oracle on
    the FULL configuration, tests/golden/make_full_size.py): 1e-12 relative, floor mean |ptw_ref|."""
    flat = np.asarray(ptw).reshape(-1)
    floor = g["mean_abs_ptw"]
    assert abs(avg - g["avg"]) <= RTOL * max(abs(g["avg"]), floor), (avg, g["avg"])
    close(flat[np.asarray(g["probe_idx"])], np.asarray(g["probe_ptw"]), floor=floor)


def test_c3_full_size(windowed):
    """C3: Frenzel-Pompe CMI 2D;2D|2D, n = 1e6, k = 4 (windowed engine)."""
    E, knn = windowed
    data, call = workloads.c3_cmi()
    subs = [(4, 5), (0, 1, 4, 5), (2, 3, 4, 5), (0, 1, 2, 3, 4, 5)]
    eps, counts = _subset_parity(E, data, call["k"], subs)
    assert (counts[3] == call["k"] - 1).all()                 # k-1 points strictly inside the k-th radius
    assert (counts[0] >= counts[1]).all() and (counts[0] >= counts[2]).all()   # z contains xz and yz
    assert (counts[1] >= counts[3]).all() and (counts[2] >= counts[3]).all()
    ptw, avg = knn.conditional_mutual_information(call["idxs_x"], call["idxs_y"], call["idxs_z"], call["k"], data)
    assert ptw.shape == (1, data.shape[1]) and np.isfinite(ptw).all()
    assert avg == pytest.approx(float(np.mean(ptw)), rel=1e-12)
    check_full_size(ptw, avg, _full_size_golden()["c3"]["cmi"])


def test_c4_full_size(windowed):
    """C4: O-information / TC / DTC / S over 8 variables, n = 250 000, k = 4: all 16 marginal count
    vectors from one joint radius, and the identities O = TC - DTC, S = TC + DTC (pointwise)."""
    E, knn = windowed
    data, call = workloads.c4_oinfo()
    m = data.shape[0]
    subs = [tuple(j for j in range(m) if j != d) for d in range(m)] + [(d,) for d in range(m)]
    eps, counts = _subset_parity(E, data, call["k"], subs)
    for d in range(m):
        assert (counts[m + d] >= counts[(d + 1) % m]).all()   # a single coordinate contains any superset of it
    tc_p, tc = knn.total_correlation(data, call["k"])
    dtc_p, dtc = knn.dual_total_correlation(data, call["k"])
    o_p, o = knn.o_information(data, call["k"])
    s_p, s = knn.s_information(data, call["k"])
    assert np.abs(o_p - (tc_p - dtc_p)).max() < 1e-9 and np.abs(s_p - (tc_p + dtc_p)).max() < 1e-9
    assert o == pytest.approx(tc - dtc, abs=1e-10) and s == pytest.approx(tc + dtc, abs=1e-10)
    g = _full_size_golden()["c4"]
    for name, (ptw, avg) in {"total_correlation": (tc_p, tc), "dual_total_correlation": (dtc_p, dtc),
                             "o_information": (o_p, o), "s_information": (s_p, s)}.items():
        check_full_size(ptw, avg, g[name])


def test_c5_full_size(windowed):
    """C5: MI(present; past) of the delay-embedded 4-channel series, n = 4e6, k = 5 (8-D joint,
    two 4-D marginals counted in their own order), and MI(x; y) = MI(y; x)."""
    E, knn = windowed
    data, call = workloads.c5_info_rate()
    subs = [(0, 1, 2, 3), (4, 5, 6, 7)]
    eps, counts = _subset_parity(E, data, call["k"], subs, n_check=192)
    assert (counts >= call["k"] - 1).all()
    ptw, avg = knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)
    ptw2, avg2 = knn.mutual_information(call["idxs_y"], call["idxs_x"], call["k"], data)
    assert np.isfinite(ptw).all() and 1.0 < avg < 1.7
    # the committed 50 000-query subset (fast oracle against the full 4e6-point set): radii and both
    # marginal count vectors bit-exact, pointwise values at 1e-12 relative
    g5 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_size_c5.npz"))
    sel = g5["sel"].astype(np.int64)
    assert np.array_equal(eps[sel], g5["eps"])
    assert np.array_equal(counts[:, sel], g5["counts"].astype(np.int64))
    close(ptw.reshape(-1)[sel], g5["ptw"])
    assert float(np.mean(ptw.reshape(-1)[sel])) == pytest.approx(_full_size_golden()["c5"]["subset_mean"], rel=1e-12)
    # swapping the groups permutes the joint coordinates: same radii, same counts; the two
    # psi terms are then subtracted in the other order, which may differ in the last bit
    assert np.abs(ptw - ptw2).max() < 1e-12 and avg == pytest.approx(avg2, rel=1e-13)


def test_heavy_tailed_data_stays_on_the_fast_path(E):
    """Outliers that stretch the range by orders of magnitude (Student-t with 1.2 degrees of
    freedom, log-normal with sigma = 3) must neither change a bit nor push the bulk of the queries
    to the float64 fallback: the fp32 copy is centred on the median and the error band is per query."""
    rng = np.random.default_rng(31)
    n = 6000
    t = rng.standard_t(1.2, size=(2, n))
    t[1] += 0.5 * t[0]
    ln = rng.lognormal(0.0, 3.0, size=(3, n))
    ln[2] *= ln[0] ** 0.25
    for rows, subs in ((t, [(0,), (1,), (0, 1)]), (ln, [(0,), (1, 2), (0, 1, 2)])):
        for key in E.stats:
            E.stats[key] = 0
        eps, counts = radius_and_counts(E, rows, 4, subs)
        ref = orc.knn_radius(rows, 4)
        assert np.array_equal(eps, ref)
        for row, s in enumerate(subs):
            assert np.array_equal(counts[row], orc.count_within(rows[list(s), :], ref)), s
        if E.config.engine != "exact":
            assert E.stats["knn_fallback_queries"] + E.stats["count_fallback_queries"] <= n // 100


def test_quantised_data_needs_no_fallback(E):
    """Dyadic-quantised / integer data (exact distance ties everywhere): the prepared set proves that
    its fp32 copy is exact, so radii and counts come straight from the fp32 filter -- bit-exact, with
    no query handed to the float64 kernels (the round-1 cliff: 86 % of the queries of 8-bit data)."""
    from oracle import ksg_oracle_c as oc

    rng = np.random.default_rng(17)
    n = 120_000
    x = rng.standard_normal(n)
    cases = {"8-bit": 8.0 / (1 << 8), "12-bit": 8.0 / (1 << 12)}
    for name, q in cases.items():
        rows = np.vstack([np.round(x / q) * q, np.round((0.6 * x + 0.8 * rng.standard_normal(n)) / q) * q,
                          np.round(rng.standard_normal(n) / q) * q])
        for key in E.stats:
            E.stats[key] = 0
        subs = [(0,), (1,), (0, 1), (1, 2), (0, 1, 2)]
        eps, counts = radius_and_counts(E, rows, 5, subs)
        sel = np.sort(rng.choice(n, size=300, replace=False))
        ref = oc.knn_radius(rows, 6, queries=rows[:, sel])
        assert np.array_equal(eps[sel], ref), name
        for row, sub in enumerate(subs):
            want = oc.count_within(rows[list(sub), :], ref, queries=rows[list(sub), :][:, sel])
            assert np.array_equal(counts[row, sel], want), (name, sub)
        if E.config.engine != "exact":
            assert E.stats["knn_fallback_queries"] == 0 and E.stats["count_fallback_queries"] == 0, (name, dict(E.stats))
    # the reference's own integer fixture (tests/test_knn.py:20-24), tiled with integer offsets
    tie = workloads.tie_fixture().astype(np.float64)
    big = np.concatenate([tie + 1000 * i for i in range(300)], axis=1)
    for key in E.stats:
        E.stats[key] = 0
    eps, counts = radius_and_counts(E, big, 1, [(0,), (1,), (2,), (0, 1)])
    ref = orc.knn_radius(big, 1)
    assert np.array_equal(eps, ref)
    for row, sub in enumerate([(0,), (1,), (2,), (0, 1)]):
        assert np.array_equal(counts[row], orc.count_within(big[list(sub), :], ref)), sub
    if E.config.engine != "exact":
        assert E.stats["knn_fallback_queries"] == 0 and E.stats["count_fallback_queries"] == 0, dict(E.stats)


def test_coordinates_beyond_the_fp32_range(E):
    """Finite float64 input farther than FLT_MAX from the median: the centred fp32 copy
    would overflow to +-inf and look like padding.  The prepared set notices (maxabs) and every query
    goes to the float64 kernels: radii and counts still equal the oracle's."""
    rng = np.random.default_rng(5)
    n = 3000
    rows = rng.standard_normal((2, n))
    rows[0, ::7] *= 1e39       # a seventh of the points lie ~1e39 away from the median
    rows[1, 5] = -3e200
    eps, counts = radius_and_counts(E, rows, 4, [(0,), (1,), (0, 1)])
    ref = orc.knn_radius(rows, 4)
    assert np.array_equal(eps, ref)
    for row, s in enumerate([(0,), (1,), (0, 1)]):
        assert np.array_equal(counts[row], orc.count_within(rows[list(s), :], ref)), s


def test_randomised_engine_agreement():
    """tools/fuzz_engines.py, a short run: random sizes / dimensions / k / subspace families /
    prepared orders / awkward distributions; windowed and dense engines == exact engine, bit for
    bit, and the public API agrees between engines (400 cases: profiles/r1ad_fuzz_400_cases.txt)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_engines.py"), "--cases", "30", "--seed", "3"],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "FUZZ_OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]


def test_rows_in_flight_prepare_equals_resident_prepare(windowed):
    """ksg_prepare_rows with per-row copy events (the sort of row d overlaps the transfer of the
    rows behind it) builds the same prepared set as ksg_prepare on resident points, for pageable
    and page-locked host arrays, and reports non-finite input in flags[0]."""
    import torch

    E, knn = windowed
    rng = np.random.default_rng(77)
    for dim, n in ((2, 200_000), (5, 70_001), (8, 30_000)):
        host = rng.standard_normal((dim, n))
        resident = E.to_device_points(host)
        torch.cuda.synchronize()
        base = E.PreparedSet(resident, E.order_for(dim))
        for src in (host, torch.from_numpy(host).pin_memory().numpy()):
            pts, ready = E.to_device_points(E.select_rows(src, range(dim)), row_events=True)
            assert len(ready) == dim
            prep = E.PreparedSet(pts, E.order_for(dim), row_ready=ready)
            assert torch.equal(prep.perm, base.perm)
            assert torch.equal(prep.sorted_f64, base.sorted_f64)
            assert int(prep.flags[0]) == 0
    bad = rng.standard_normal((3, 5000))
    for value in (np.nan, -np.nan, np.inf, -np.inf):
        rows = bad.copy()
        rows[1, 1234] = value
        prep = E.PreparedSet(E.to_device_points(rows), -1)
        assert int(prep.flags[0]) == 1, value
    a = rng.standard_normal((2, 150_000))
    a[1] += a[0]
    pinned = torch.from_numpy(a).pin_memory().numpy()
    p1, m1 = knn.mutual_information((0,), (1,), 4, a)
    p2, m2 = knn.mutual_information((0,), (1,), 4, pinned)
    assert np.array_equal(p1, p2) and m1 == m2


def test_speculative_run_recomputes_when_the_filter_flags_queries(windowed):
    """MI of two scalar series runs speculatively (counts and psi enqueued behind the k-NN pass,
    its flag count read with the result).  On tie-heavy data the filter flags queries: the settled
    result must equal the oracle's pointwise values; on continuous data nothing is flagged and the
    speculative values stand."""
    E, knn = windowed
    rng = np.random.default_rng(12)
    n = 6000
    x = rng.standard_normal(n)
    y = 0.6 * x + 0.8 * rng.standard_normal(n)
    # (decimal quantisation: multiples of 0.03 are not exact in fp32, so the tie-saturated lists cannot be
    #  decided by the filter; dyadic quantisation is -- the prepared set proves its fp32 copy exact -- and
    #  then nothing is escalated, test_quantised_data_needs_no_fallback)
    for rows, expect_escalation in ((np.round(np.vstack([x, y]) / 0.03) * 0.03, True),
                                    (np.round(np.vstack([x, y]) * 32) / 32, False), (np.vstack([x, y]), False)):
        for key in E.stats:
            E.stats[key] = 0
        with np.errstate(all="ignore"):
            ptw, avg = knn.mutual_information((0,), (1,), 4, rows)
            ref_ptw, ref_avg = orc.mutual_information((0,), (1,), 4, rows)
        close_pair(ptw, ref_ptw, avg, ref_avg)
        escalated = (E.stats["knn_fallback_queries"] + E.stats["knn_deferred_queries"]
                     + E.stats["knn_wide_margin_passes"]) > 0
        assert escalated == expect_escalation
        # the device-resident entry points (bench.py's step) settle the same way
        import torch

        from syntropy_b200.knn.shannon import _xy_masks, mi1_program

        pts = E.to_device_points(rows)
        status = E.Status(pts.device)
        pw = E.ksg1_device(pts, 4, _xy_masks((0,), (1,)), mi1_program, status=status)
        assert pw.redo is not None
        full, total = E.finish_device(pw, n, status=status)
        pw, full, total = E.settle(pw, n, status, full, total)
        torch.cuda.synchronize()
        close(full.cpu().numpy(), np.asarray(ref_ptw).reshape(-1))


# ------------------------------------------------------------------ syntropy.mixed, knn branch
def test_mixed_knn_branch_vs_reference_golden(E, golden):
    """shannon_entropy / conditional_entropy / mutual_information with continuous_estimator="knn"
    against the unmodified reference's outputs (tests/golden/mixed_knn.npz): per-class radii are
    exact, so the pointwise arrays agree to 1e-12 (psi / log ulps), averages likewise."""
    from syntropy_b200 import mixed

    g = golden("mixed_knn")
    with np.errstate(all="ignore"):
        for name, (d, c, k) in workloads.mixed_cases().items():
            ptw, avg = mixed.shannon_entropy(d, c, continuous_estimator="knn", k=k)
            close_pair(ptw, g[f"{name}_h_ptw"], avg, g[f"{name}_h_avg"])
            for cond in ("discrete", "continuous"):
                ptw, avg = mixed.conditional_entropy(d, c, conditional=cond, continuous_estimator="knn", k=k)
                close_pair(ptw, g[f"{name}_ce_{cond}_ptw"], avg, g[f"{name}_ce_{cond}_avg"])
            ptw, avg = mixed.mutual_information(d, c, continuous_estimator="knn", k=k)
            assert ptw.shape == (1, d.shape[1])
            close_pair(ptw, g[f"{name}_mi_ptw"], avg, g[f"{name}_mi_avg"])


def test_mixed_knn_properties_at_the_reference_test_size(windowed):
    """The reference's own statistical checks (tests/test_mixed.py:279-324, n = 50 000, k = 5,
    tolerance 2e-2): well-separated classes saturate I(D;C) at ln 2, independent C gives 0, the
    pointwise mean equals the average; plus the chain rule H(D,C) = H(D) + H(C|D) exactly."""
    from syntropy_b200 import mixed

    rng = np.random.default_rng(0)
    n = 50_000
    d = rng.integers(0, 2, size=(1, n))
    c = rng.standard_normal((1, n)) + 5.0 * d
    ptw, mi = mixed.mutual_information(d, c, continuous_estimator="knn", k=5)
    assert mi == pytest.approx(np.log(2.0), abs=2e-2)
    assert ptw.mean() == pytest.approx(mi, abs=1e-9)
    c_ind = rng.standard_normal((1, n))
    assert mixed.mutual_information(d, c_ind, continuous_estimator="knn", k=5)[1] == pytest.approx(0.0, abs=2e-2)
    _, joint = mixed.shannon_entropy(d, c, continuous_estimator="knn", k=5)
    _, h_c_given_d = mixed.conditional_entropy(d, c, conditional="discrete", continuous_estimator="knn", k=5)
    p = np.bincount(d[0]) / n
    assert joint == pytest.approx(-(p * np.log(p)).sum() + h_c_given_d, abs=1e-12)
    # the per-class calls are knn.differential_entropy on the class's columns
    from syntropy_b200 import knn

    ptw_h, _ = mixed.shannon_entropy(d, c, continuous_estimator="knn", k=5)
    for state in (0, 1):
        mask = d[0] == state
        ptw_c, _ = knn.differential_entropy(c[:, mask], k=5)
        assert np.array_equal(ptw_h[0, mask], (0.0 - np.log(p[state])) + ptw_c[0])
