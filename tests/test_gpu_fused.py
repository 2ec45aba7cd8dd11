"""GPU checks of the round-2 launch fusions and of the warp-level k-NN kernel against the kernels they
replace (all through the C ABI):

* ``gather_tile_kernel`` (gather + shadow + inverse permutation + boxes + exactness flag, one launch) builds
  prepared sets identical, byte for byte, to the four separate kernels (``KSG_B200_FUSED_GATHER=0``);
* ``knn_wunits_kernel`` (every warp its own worker, unsorted lists) returns the radii, flags and neighbour
  lists of ``knn_units_kernel`` (``KSG_B200_UNITS=cta``) -- continuous, quantised, heavy-tailed data, 1 to 4
  coordinates, short and long lists, cross-set queries;
* ``ksg_axis_count_multi`` == one ``ksg_axis_count`` per coordinate;
* ``ksg_psi_finish`` / ``ksg_scatter_sum_f64`` == ``ksg_psi_epilogue`` + ``ksg_scatter_f64`` (+ ``ksg_sum_f64``
  up to the summation order; the ticket comes back to 0).
"""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sets():
    from syntropy_b200 import workloads

    rng = np.random.default_rng(19)
    heavy = rng.standard_t(1.3, size=(2, 60_000))
    return [("c2", workloads.c2_mi_1m(n=300_000)[0]),
            ("c3", workloads.c3_cmi(n=40_000)[0]),
            ("3d", rng.standard_normal((3, 5_000))),
            ("1d", rng.standard_normal((1, 7_001))),
            ("4d", rng.standard_normal((4, 33_000)) + 1e3),
            ("quantised", np.round(rng.standard_normal((2, 80_000)) * 8) / 8 + 0.0),
            ("integers", rng.integers(-50, 50, size=(3, 20_000)).astype(np.float64)),
            ("heavy", heavy),
            ("8d", rng.standard_normal((8, 9_000))),
            ("tiny", rng.standard_normal((2, 37)))]


def _prepared_digest():
    """sha1 over everything ksg_prepare writes (orders, both copies, boxes, flags) for a few point sets."""
    import torch

    from syntropy_b200 import _engine as E

    out = []
    for name, rows in _sets():
        pts = E.to_device_points(rows)
        prep = E.PreparedSet(pts, E.order_for(rows.shape[0]))
        E._cabi.check(E._cabi.load().ksg_prepare_join(C.byref(prep.c), E._stream()), "ksg_prepare_join")
        torch.cuda.synchronize()
        h = hashlib.sha1()
        dim, n, ld, dp, npad = prep.dim, prep.n, prep.ld, int(prep.c.dim_padded), int(prep.c.n_padded)
        tiles = npad // 512
        h.update(prep.perm.cpu().numpy().tobytes())
        h.update(prep.sorted_f64.cpu().numpy().tobytes())
        h.update(prep._view(prep.c.shadow_f32, npad * dp * 4, torch.float32).cpu().numpy().tobytes())
        h.update(prep._view(prep.c.tile_box, tiles * dp * 2 * 4, torch.float32).cpu().numpy().tobytes())
        h.update(prep._view(prep.c.sub_box, tiles * 16 * dp * 2 * 4, torch.float32).cpu().numpy().tobytes())
        h.update(prep._view(prep.c.axis_vals, dim * ld * 8, torch.float64).view(dim, ld)[:, :n].cpu().numpy().tobytes())
        h.update(prep._view(prep.c.axis_pos, dim * ld * 4, torch.int32).view(dim, ld)[:, :n].cpu().numpy().tobytes())
        flags = prep._view(prep.c.flags, 64, torch.int32).cpu().numpy()
        h.update(flags[[0, 8]].tobytes())  # non-finite input, "the fp32 copy is exact"
        out.append("%s:%s:exact32=%d" % (name, h.hexdigest(), int(flags[8])))
    return "|".join(out)


def _knn_digest_high():
    return _knn_digest(high=True)


def _knn_digest(high=False):
    """sha1 over radii / flags / neighbour lists of the windowed k-NN stage (first pass, nothing escalated);
    ``high``: the sets of 5 to 8 coordinates instead of those of 1 to 4."""
    import torch

    from syntropy_b200 import _engine as E

    out = []
    rng = np.random.default_rng(23)
    sets = _sets()
    if high:
        r2 = np.random.default_rng(29)
        sets = [s for s in sets if s[1].shape[0] > 4] + [
            ("5d", r2.standard_normal((5, 30_000))), ("7d heavy", r2.standard_t(2.0, size=(7, 12_000))),
            ("6d quantised", np.round(r2.standard_normal((6, 20_000)) * 4) / 4),
            ("8d clusters", np.concatenate([r2.standard_normal((8, 4_000)) * 0.1 + 3 * c for c in range(4)], axis=1))]
    for name, rows in sets:
        dim = rows.shape[0]
        if (dim > 4) != high:
            continue
        pts = E.to_device_points(rows)
        prep = E.PreparedSet(pts, E.order_for(dim) if high else -1)
        n = prep.n
        # (k' + margin = 8, 12, 16, 24: the list capacities of the warp-level kernel, so that both kernels keep
        #  lists of the same length and flag the same tie-saturated queries; 44 and 36: too long for it)
        for kprime, margin, want_idx in ((2, 6, False), (6, 6, False), (6, 6, True), (9, 7, False), (12, 12, False),
                                         (20, 24, False), (30, 6, False)):
            if kprime >= n:
                continue
            eps, idx, flags, nflag = E._fast_knn_call(prep, kprime, 0, n, True, want_indices=want_idx, margin=margin)
            torch.cuda.synchronize()
            h = hashlib.sha1()
            f = flags.cpu().numpy()
            e = eps.cpu().numpy()
            h.update(f.tobytes())
            h.update(e[f == 0].tobytes())
            if want_idx:
                h.update(idx.cpu().numpy()[f == 0].tobytes())
            out.append("%s/k%d/m%d:%s:flagged=%d" % (name, kprime, margin, h.hexdigest(), int(nflag.item())))
        # a query slice that starts inside a block, and a cross-set search
        q0, q1 = n // 3 + 5, n // 3 + 5 + min(n // 2, 4_321)
        eps, _, flags, _ = E._fast_knn_call(prep, 4, q0, q1, True)
        out.append("%s/slice:%s" % (name, hashlib.sha1(eps.cpu().numpy().tobytes() + flags.cpu().numpy().tobytes()).hexdigest()))
        if dim >= 2 and n >= 1000 and int(prep.c.primary) == -1:
            other = torch.from_numpy(np.ascontiguousarray(rows[:, ::-1] * 1.05 + 0.01 * rng.standard_normal(rows.shape))).cuda()
            cross = E.CrossSet(prep, other)
            r = E.fast_knn_cross(cross, 3, 0, other.shape[1])
            out.append("%s/cross:%s" % (name, hashlib.sha1(r.cpu().numpy().tobytes()).hexdigest()))
    return "|".join(out)


def _in_subprocess(fn_name, **env):
    code = ("import sys; sys.path.insert(0, %r); from tests import test_gpu_fused as t; print(t.%s())" % (ROOT, fn_name))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env),
                         timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    return out.stdout.strip().splitlines()[-1]


def _diff(a, b):
    return [x + "  !=  " + y for x, y in zip(a.split("|"), b.split("|")) if x != y]


def test_fused_gather_builds_identical_prepared_sets():
    mine = _prepared_digest()
    assert "quantised:" in mine and "exact32=1" in mine and "exact32=0" in mine  # both outcomes of the exactness fold
    ref = _in_subprocess("_prepared_digest", KSG_B200_FUSED_GATHER="0")
    assert mine == ref, _diff(mine, ref)


def test_warp_level_knn_equals_the_cta_level_kernel():
    mine = _knn_digest()
    ref = _in_subprocess("_knn_digest", KSG_B200_UNITS="cta")
    assert mine == ref, _diff(mine, ref)


def _count_digest():
    """sha1 over the raw marginal counts of the windowed count kernels: all leave-one-out subspaces of 5- to 8-D sets
    (fused pass) and the full space of 2- to 4-D sets (own-order pass), for radii taken from the k-NN stage,
    whole sets and a query slice that starts inside a block."""
    import torch

    from syntropy_b200 import _engine as E

    out = []
    r2 = np.random.default_rng(31)
    sets = [("5d", r2.standard_normal((5, 9_000))), ("6d quantised", np.round(r2.standard_normal((6, 7_000)) * 4) / 4),
            ("7d heavy", r2.standard_t(2.0, size=(7, 6_000))), ("8d", r2.standard_normal((8, 9_000))),
            ("8d clusters", np.concatenate([r2.standard_normal((8, 2_500)) * 0.1 + 3 * c for c in range(4)], axis=1)),
            ("8d duplicates", np.repeat(r2.standard_normal((8, 1_500)), 3, axis=1)),
            ("2d", r2.standard_normal((2, 30_000))), ("3d heavy", r2.standard_t(1.5, size=(3, 20_000))),
            ("4d", r2.standard_normal((4, 40_000))), ("4d quantised", np.round(r2.standard_normal((4, 15_000)) * 8) / 8),
            ("4d small", r2.standard_normal((4, 300)))]
    for name, rows in sets:
        dim, n = rows.shape
        pts = E.to_device_points(rows)
        prep = E.PreparedSet(pts, E.order_for(dim))
        eps = E.fast_knn(prep, 5, 0, n, True)
        full = (1 << dim) - 1
        masks = [full & ~(1 << s) for s in range(dim)] if dim >= 5 else [full]
        for q0, q1 in ((0, n), (n // 3 + 5, n // 3 + 5 + min(n // 2, 4_321))):
            for strict in (True, False):
                c = E.fast_counts(prep, masks, eps[q0:q1].contiguous(), q0, q1, True, strict=strict)
                torch.cuda.synchronize()
                out.append("%s/%d-%d/%d:%s" % (name, q0, q1, strict, hashlib.sha1(c.cpu().numpy().tobytes()).hexdigest()))
    return "|".join(out)


def test_grouped_count_kernels_equal_the_whole_warp_and_block_level_kernels():
    """8-query groups (own-order marginal pass, fused leave-one-out pass with sub-tile culling on the second-largest
    box gap) against the kernels they replace; the leave-one-out kernel is forced on for these small sets."""
    mine = _in_subprocess("_count_digest", KSG_B200_LOO_GROUPS="force")
    ref = _in_subprocess("_count_digest", KSG_B200_LOO_GROUPS="0", KSG_B200_COUNT_GROUPS="0")
    assert mine == ref, _diff(mine, ref)
    eight = _in_subprocess("_count_digest", KSG_B200_LOO_GROUPS="0", KSG_B200_COUNT_GROUPS="8")
    assert eight == ref, _diff(eight, ref)


def test_grouped_knn_equals_the_whole_warp_kernel():
    """knn_gunits_kernel (8-query groups inside the filter warps, 5 to 8 coordinates) == knn_units_kernel."""
    mine = _knn_digest_high()
    assert mine.count("|") > 20
    ref = _in_subprocess("_knn_digest_high", KSG_B200_GROUPS="0")
    assert mine == ref, _diff(mine, ref)


def test_axis_count_multi_equals_single_launches():
    import torch

    from syntropy_b200 import _engine as E

    lib = E._cabi.load()
    rng = np.random.default_rng(5)
    for rows, k in ((rng.standard_normal((3, 50_000)), 4), (np.round(rng.standard_normal((2, 20_000)) * 4) / 4, 3)):
        pts = E.to_device_points(rows)
        prep = E.PreparedSet(pts, -1)
        n, dim = prep.n, prep.dim
        eps = E.fast_knn(prep, k + 1, 0, n, True)
        for q0, q1 in ((0, n), (n // 4 + 3, n // 2 + 11)):
            nq = q1 - q0
            e = eps[q0:q1].contiguous()
            axes = list(range(dim))[::-1]
            multi = torch.empty((dim, nq), dtype=torch.int32, device=pts.device)
            arr = (C.c_int32 * dim)(*axes)
            E._cabi.check(lib.ksg_axis_count_multi(C.byref(prep.c), arr, dim, q0, q1, E._ptr(e), 1, -1, E._ptr(multi),
                                                   E._stream()), "ksg_axis_count_multi")
            for row, axis in enumerate(axes):
                one = torch.empty(nq, dtype=torch.int32, device=pts.device)
                E._cabi.check(lib.ksg_axis_count(C.byref(prep.c), axis, q0, q1, E._ptr(e), 1, -1, E._ptr(one), E._stream()),
                              "ksg_axis_count")
                assert torch.equal(multi[row], one)


def test_fused_tail_equals_epilogue_scatter_sum():
    import torch

    from syntropy_b200 import _engine as E

    lib = E._cabi.load()
    rng = np.random.default_rng(6)
    for n in (1, 37, 1000, 300_001):
        dev = torch.device("cuda")
        counts = torch.from_numpy(rng.integers(1, max(n, 2), size=(3, n)).astype(np.int32)).to(dev)  # (psi(0) = -inf)
        perm = torch.from_numpy(rng.permutation(n).astype(np.int32)).to(dev)
        prog = E.PsiProgram(init=0.25, psi_n=E.psi_host(n), scale=0.5)
        prog.add(0, -1)
        prog.add(1, -1, count_offset=0)
        prog.add(2, +1, centered=True, divisor=3.0)
        ptw = E.psi_epilogue(counts, prog, n)
        ref = torch.empty_like(ptw)
        E._cabi.check(lib.ksg_scatter_f64(E._ptr(ptw), E._ptr(perm), n, E._ptr(ref), E._stream()), "ksg_scatter_f64")
        ref_sum = float(E.device_sum(ref).item())
        table = E.psi_table(n + 2)
        steps = (E._cabi.PsiStep * len(prog.steps))(*[E._cabi.PsiStep(*st) for st in prog.steps])
        ws = torch.empty(int(lib.ksg_finish_workspace(n)), dtype=torch.uint8, device=dev)
        status = E.Status(dev)
        for _ in range(2):  # the ticket must come back to 0: the second call works like the first
            out = torch.full((n,), -7.0, dtype=torch.float64, device=dev)
            E._cabi.check(lib.ksg_psi_finish(E._ptr(counts), n, 3, steps, len(prog.steps), prog.init, prog.psi_n, 1, prog.scale,
                                             E._ptr(table), table.numel(), E._ptr(perm), E._ptr(out), E._ptr(status.total),
                                             E._ptr(status.ticket), E._ptr(ws), ws.numel(), E._stream()), "ksg_psi_finish")
            assert torch.equal(out, ref)
            got = float(status.total.item())
            assert abs(got - ref_sum) <= 1e-12 * max(abs(ref_sum), float(ref.abs().sum().item()) * 1e-3 + 1e-300)
            assert int(status.ticket.item()) == 0
        out2 = torch.full((n,), -7.0, dtype=torch.float64, device=dev)
        total2 = torch.zeros(1, dtype=torch.float64, device=dev)
        E._cabi.check(lib.ksg_scatter_sum_f64(E._ptr(ptw), E._ptr(perm), n, E._ptr(out2), E._ptr(total2), E._ptr(status.ticket),
                                              E._ptr(ws), ws.numel(), E._stream()), "ksg_scatter_sum_f64")
        assert torch.equal(out2, ref)
        assert float(total2.item()) == got  # the same tree over the same prepared order
        assert int(status.ticket.item()) == 0


def test_public_call_is_the_same_with_and_without_the_fused_tail():
    from syntropy_b200 import _engine as E
    from syntropy_b200 import knn, workloads

    data, call = workloads.c2_mi_1m(n=200_000)
    ptw1, avg1 = knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)
    E.config.fused_tail = False
    try:
        ptw0, avg0 = knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)
    finally:
        E.config.fused_tail = True
    assert np.array_equal(ptw1, ptw0)
    assert abs(avg1 - avg0) <= 1e-13 * abs(avg0)


def test_bucketed_axis_counts_equal_the_sorted_ones():
    """ksg_axis_count_buckets on a set with lazy axes (primary = -3) == ksg_axis_count_multi on the ordinary
    set; tied data raises the flag instead; ksg_prepare_axes then turns the lazy set into an ordinary one."""
    import torch

    from syntropy_b200 import _engine as E

    lib = E._cabi.load()
    rng = np.random.default_rng(31)
    cases = [("normal", rng.standard_normal((2, 200_000)), False),
             ("3d offset", rng.standard_normal((3, 50_000)) * 1e-6 + 1e3, False),
             ("heavy", rng.standard_t(1.2, size=(2, 60_000)), False),
             ("tiny", rng.standard_normal((2, 41)), False),
             ("1d", rng.standard_normal((1, 9_000)), False),
             ("decimal grid", np.round(rng.standard_normal((2, 100_000)), 2), None),   # few hundred values per level: may go either way
             ("quantised", np.round(rng.standard_normal((2, 300_000)) * 8) / 8, True)]
    for name, rows, expect_flag in cases:
        pts = E.to_device_points(rows)
        dim, n = rows.shape
        ref = E.PreparedSet(pts, -1)
        lazy = E.PreparedSet(pts, -3)
        assert int(lazy.c.axes_mode) > 0 and int(ref.c.axes_mode) == 0, name
        assert torch.equal(lazy.perm, ref.perm), name   # the same curve order
        k = 4
        eps = E.fast_knn(ref, k + 1, 0, n, True)
        eps_lazy = E.fast_knn(lazy, k + 1, 0, n, True)   # (the k-NN stage never needs the axes)
        assert torch.equal(eps, eps_lazy), name
        eps[::17] = float("inf")
        eps[5::23] = 0.0
        axes = (C.c_int32 * dim)(*range(dim))
        for q0, q1 in ((0, n), (n // 3 + 1, n // 2 + 7)):
            nq = q1 - q0
            e = eps[q0:q1].contiguous()
            want = torch.empty((dim, nq), dtype=torch.int32, device=pts.device)
            E._cabi.check(lib.ksg_axis_count_multi(C.byref(ref.c), axes, dim, q0, q1, E._ptr(e), 1, -1, E._ptr(want), E._stream()),
                          "ksg_axis_count_multi")
            for strict in (1, 0):
                if not strict:
                    E._cabi.check(lib.ksg_axis_count_multi(C.byref(ref.c), axes, dim, q0, q1, E._ptr(e), 0, 0, E._ptr(want),
                                                           E._stream()), "ksg_axis_count_multi")
                got = torch.full((dim, nq), -9, dtype=torch.int32, device=pts.device)
                nflag = torch.zeros(1, dtype=torch.int32, device=pts.device)
                E._cabi.check(lib.ksg_axis_count_buckets(C.byref(lazy.c), axes, dim, q0, q1, E._ptr(e), strict, -1 if strict else 0,
                                                         E._ptr(got), E._ptr(nflag), E._stream()), "ksg_axis_count_buckets")
                flagged = int(nflag.item()) != 0
                if expect_flag is not None:
                    assert flagged == expect_flag, (name, int(nflag.item()))
                if not flagged:
                    assert torch.equal(got, want), (name, q0, strict, int((got != want).sum().item()))
        # exact entry points refuse a lazy set ...
        one = torch.empty(n, dtype=torch.int32, device=pts.device)
        assert lib.ksg_axis_count(C.byref(lazy.c), 0, 0, n, E._ptr(eps), 1, -1, E._ptr(one), E._stream()) != 0
        # ... until its axes are built: then it counts like the ordinary set
        lazy.ensure_axes()
        assert int(lazy.c.axes_mode) == 0
        got = E.fast_counts(lazy, [1 << d for d in range(dim)], eps, 0, n, True)
        want = E.fast_counts(ref, [1 << d for d in range(dim)], eps, 0, n, True)
        assert torch.equal(got, want), name


def test_public_mi_with_and_without_lazy_axes():
    from syntropy_b200 import _engine as E
    from syntropy_b200 import knn, workloads

    rng = np.random.default_rng(3)
    sets = [workloads.c2_mi_1m(n=150_000)[0], np.round(rng.standard_normal((2, 120_000)) * 8) / 8,
            rng.integers(0, 9, size=(2, 3_000)).astype(np.float64)]
    built = []
    for data in sets:
        for key in E.stats:
            E.stats[key] = 0
        ptw1, avg1 = knn.mutual_information((0,), (1,), 4, data)
        built.append(E.stats["lazy_axes_built"])
        E.config.lazy_axes = False
        try:
            ptw0, avg0 = knn.mutual_information((0,), (1,), 4, data)
        finally:
            E.config.lazy_axes = True
        assert np.array_equal(ptw1, ptw0, equal_nan=True)
        assert avg1 == avg0 or abs(avg1 - avg0) <= 1e-13 * abs(avg0) or (np.isnan(avg1) and np.isnan(avg0)) or avg1 == avg0
    assert built[0] == 0    # continuous data: settled from the value buckets, no coordinate was ever sorted
    assert built[1] >= 1    # tied data went through the give-up path (flag -> ksg_prepare_axes -> sorted counts)
