"""CPU-side checks of the host orchestration (no CUDA calls): engine / ordering decisions, which
marginals get their own prepared set, the shared-memory guard of the fast path, the lazy row
selection, the ctypes mirrors of the header structs and the workspace arithmetic of the C ABI."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def E():
    from syntropy_b200 import _engine

    saved = {k: getattr(_engine.config, k) for k in ("engine", "primary", "kd_min_dim", "own_order", "own_order_max_dim")}
    yield _engine
    for k, v in saved.items():
        setattr(_engine.config, k, v)


def test_masks_and_dims_roundtrip(E):
    assert E.mask_of((0, 2, 5)) == 0b100101
    assert E._dims_of(0b100101) == [0, 2, 5]
    assert E._dims_of(E.mask_of(range(8))) == list(range(8))


def test_prepared_order_selection(E):
    E.config.primary, E.config.kd_min_dim = -1, 7
    assert [E.order_for(d) for d in (1, 2, 4, 6, 7, 8)] == [-1, -1, -1, -1, -2, -2]
    E.config.primary = -2
    assert E.order_for(2) == -2
    E.config.primary = 3                       # a fixed coordinate is clipped to the set's dimension
    assert E.order_for(2) == 1 and E.order_for(8) == 3


def test_which_marginals_get_their_own_prepared_set(E):
    E.config.engine, E.config.own_order, E.config.own_order_max_dim = "windowed", True, 4
    cmi = [E.mask_of((4, 5)), E.mask_of((0, 1, 4, 5)), E.mask_of((2, 3, 4, 5))]          # C3: z, xz, yz of 6-D
    assert E.own_order_subspaces(cmi, 6) == [0, 1, 2]
    loo8 = [E.mask_of(j for j in range(8) if j != d) for d in range(8)] + [1 << d for d in range(8)]
    assert E.own_order_subspaces(loo8, 8) == []                  # 7-D leave-one-out: fused; 1-D: binary search
    assert E.own_order_subspaces([0b01, 0b10, 0b11], 2) == []    # the joint space itself stays in the joint set
    assert E.own_order_subspaces([E.mask_of(range(4)), E.mask_of(range(4, 8))], 8) == [0, 1]   # C5
    E.config.engine = "dense"
    assert E.own_order_subspaces(cmi, 6) == []
    E.config.engine, E.config.own_order = "windowed", False
    assert E.own_order_subspaces(cmi, 6) == []


def test_fast_path_guard_by_dimension_and_list_size(E):
    E.config.engine = "windowed"
    assert E._use_fast(2, 6) and E._use_fast(8, 6)
    assert not E._use_fast(9, 6)                                  # more than 8 joint coordinates
    assert E._use_fast(2, 40) and not E._use_fast(2, 50)          # 512 queries per block: lists must fit in smem
    assert E._use_fast(8, 60) and not E._use_fast(8, 62)          # k' + 3 <= 64
    E.config.engine = "exact"
    assert not E._use_fast(2, 6)
    E.config.engine = "nonsense"
    with pytest.raises(ValueError):
        E._use_fast(2, 6)


def test_row_selection_is_lazy_and_mirrors_fancy_indexing(E):
    data = np.arange(40, dtype=np.int64).reshape(5, 8)
    sel = E.select_rows(data, (3, 0, -1))
    assert sel.shape == (3, 8)
    assert np.array_equal(np.stack([sel.row(r) for r in range(3)]), data[(3, 0, -1), :])
    assert sel.row(0).base is not None                            # a view, not a copy
    with pytest.raises(ValueError):
        E.select_rows(np.zeros(7), (0,))
    assert not E._rows_are_page_locked(E.select_rows(np.zeros((2, 100)), (0, 1)), 2)   # small / pageable


def test_host_copy_casts_like_astype(E):
    src = (np.arange(700_000) % 251).astype(np.int32)
    dst = np.empty(src.size, dtype=np.float64)
    E._host_copy(dst, src)                                         # split over the copy threads
    assert np.array_equal(dst, src.astype(np.float64))
    small = np.empty(5)
    E._host_copy(small, np.array([1, 2, 3, 4, 5], dtype=np.uint8))
    assert small.tolist() == [1.0, 2.0, 3.0, 4.0, 5.0]


def _header_struct_fields(name):
    text = open(os.path.join(ROOT, "include", "ksg_b200.h")).read()
    body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\}\s*" + name + r"\s*;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [re.search(r"(\w+)\s*$", decl.strip()).group(1) for decl in body.split(";") if decl.strip()]


def test_ctypes_structs_mirror_the_header():
    from syntropy_b200 import _cabi

    assert [f[0] for f in _cabi.Prepared._fields_] == _header_struct_fields("ksg_prepared")
    assert [f[0] for f in _cabi.Cross._fields_] == _header_struct_fields("ksg_cross")
    # all-pointer tails and 8-byte scalars: no padding surprises
    assert C.sizeof(_cabi.Prepared) == 8 + 4 * 4 + 8 * 2 + 8 * 13
    assert C.sizeof(_cabi.Cross) == 8 * 7


def test_workspace_arithmetic_without_a_device():
    from syntropy_b200 import _cabi

    lib = _cabi.load()
    n = 1_000_000
    assert lib.ksg_prepare_workspace(n, 2) > n * (4 + 16 + 8 + 24)
    assert lib.ksg_prepare_workspace(0, 2) == 0 and lib.ksg_prepare_workspace(n, 33) == 0
    w_full = lib.ksg_fast_knn_workspace(n, n, 2, 6)
    w_slice = lib.ksg_fast_knn_workspace(n, n // 8, 2, 6)
    assert 0 < w_slice < w_full
    assert lib.ksg_fast_count_workspace(n, n, 4, 1) >= n * 4
    assert lib.ksg_cross_prepare_workspace(n, n, 3) > n * (4 + 24 + 16 + 4)
    # null pointers are refused with a status code and a message, not a crash
    assert lib.ksg_fast_knn_cross(None, None, 0, 0, 1, None, None, None, None, 0, None) == -1
    assert b"ksg_fast_knn_cross" in lib.ksg_last_error()
    assert lib.ksg_sum_i32(None, 5, None, None) == -1


def test_int32_counts_gather_like_float64_slices():
    """dist.gather_slices is dtype-agnostic (the own-order path gathers int32 count vectors)."""
    import torch

    from syntropy_b200 import dist

    dist.configure("off")
    try:
        v = torch.arange(10, dtype=torch.int32)
        assert dist.gather_slices(v, 10) is v and dist.my_slice(10) == (0, 10)
    finally:
        dist.configure("auto")


def test_mixed_class_partition_equals_the_reference_masks():
    """syntropy_b200.mixed groups samples by a counted mixed-radix key (or np.unique's inverse); the
    reference builds one boolean mask per distinct column (mixed/shannon.py:53-65): same classes,
    same order, same members -- for small-range integers (fast path), negative values, several
    discrete variables, wide-range integers and floats (general path)."""
    from syntropy_b200 import workloads
    from syntropy_b200.mixed.shannon import _classes

    rng = np.random.default_rng(8)
    cases = {name: d for name, (d, _, _) in workloads.mixed_cases().items()}
    cases["negative"] = rng.integers(-3, 2, size=(2, 500))
    cases["three vars, int8"] = rng.integers(0, 3, size=(3, 700)).astype(np.int8)
    cases["bool"] = rng.integers(0, 2, size=(2, 300)).astype(bool)
    cases["wide range"] = rng.choice(np.array([-2**40, 5, 2**50]), size=(2, 400))
    cases["float labels"] = rng.choice(np.array([0.5, 1.5, -2.0]), size=(1, 300))
    cases["one class"] = np.full((1, 50), 3)
    for name, d in cases.items():
        probs, class_id, order, bounds = _classes(d)
        unq, counts = np.unique(d, axis=1, return_counts=True)
        assert np.array_equal(probs, counts / counts.sum()), name
        assert bounds.shape[0] == unq.shape[1] + 1 and bounds[-1] == d.shape[1]
        for i in range(unq.shape[1]):
            mask = np.all(d == unq[:, [i]], axis=0)
            assert np.array_equal(order[bounds[i]:bounds[i + 1]], np.flatnonzero(mask)), (name, i)
            assert np.all(class_id[mask] == i), (name, i)


def test_mixed_rejects_what_it_does_not_implement():
    from syntropy_b200 import mixed

    d = np.zeros((1, 10), dtype=np.int64)
    c = np.zeros((1, 10))
    with pytest.raises(NotImplementedError):
        mixed.mutual_information(d, c)  # default estimator is the Gaussian plug-in
    with pytest.raises(AssertionError):
        mixed.shannon_entropy(d, c, continuous_estimator="kde")
    with pytest.raises(AssertionError):
        mixed.conditional_entropy(d, c[:, :5], continuous_estimator="knn")
    with pytest.raises(AssertionError):
        mixed.conditional_entropy(d, c, conditional="both", continuous_estimator="knn")
