"""The key sorts behind ksg_prepare on the GPU: ksg_sort_f64 (sample sort, csrc/sample_sort.cuh)
against numpy's stable argsort and against the radix sort ksg_prepare uses, and prepared sets that
are identical under both."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_sort(keys, algorithm=0):
    import torch

    from syntropy_b200 import _cabi
    from syntropy_b200 import _engine as E

    lib = _cabi.load()
    dev = E.require_cuda()
    k = torch.from_numpy(np.ascontiguousarray(keys, dtype=np.float64)).to(dev)
    n = k.numel()
    out_k = torch.empty_like(k)
    out_i = torch.empty(n, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.ksg_sort_workspace(n)), dtype=torch.uint8, device=dev)
    rc = lib.ksg_sort_f64(C.c_void_p(k.data_ptr()), n, C.c_void_p(out_k.data_ptr()), C.c_void_p(out_i.data_ptr()),
                          algorithm, C.c_void_p(ws.data_ptr()), ws.numel(), E._stream())
    _cabi.check(rc, "ksg_sort_f64")
    torch.cuda.synchronize()
    return out_k.cpu().numpy(), out_i.cpu().numpy()


def cases():
    rng = np.random.default_rng(3)
    yield "one", np.array([2.5])
    yield "five", rng.standard_normal(5)
    for n in (4096, 4097, 5000, 100_003, 1_000_000, (1 << 21), (1 << 21) + 7):
        yield f"normal_{n}", rng.standard_normal(n)
    yield "all_equal", np.full(300_000, 1.25)
    yield "sorted", np.sort(rng.standard_normal(250_000))
    yield "reversed", np.sort(rng.standard_normal(250_000))[::-1].copy()
    yield "few_values", rng.integers(0, 7, size=400_000).astype(np.float64)
    yield "periodic", np.tile(np.arange(1024, dtype=np.float64), 300)
    z = rng.standard_normal(200_000)
    z[::5] = 0.0
    z[1::7] = -0.0
    z[3::1001] = np.inf
    z[4::1003] = -np.inf
    yield "zeros_infs", z
    yield "heavy_tail", rng.standard_t(1.1, size=500_000) * 1e30
    yield "tiny", rng.standard_normal(300_000) * 1e-300


def test_sort_matches_stable_argsort():
    for name, keys in cases():
        got_k, got_i = gpu_sort(keys)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(got_k, keys[order]), name
        same = got_i == order
        if not same.all():
            # the only freedom: -0.0 and +0.0 compare equal for numpy but are distinct keys here
            bad = np.flatnonzero(~same)
            assert np.all(keys[got_i[bad]] == 0.0) and np.all(keys[order[bad]] == 0.0), name
        assert np.array_equal(np.sort(got_i), np.arange(keys.size)), name
        if keys.size <= (1 << 21) and not (keys == 0.0).any():
            ref_k, ref_i = gpu_sort(keys, algorithm=1)   # cub::DeviceRadixSort, the path it replaces
            assert np.array_equal(got_k, ref_k) and np.array_equal(got_i, ref_i), name


def _prepared_digest():
    """sha1 over perm, axis_vals / axis_pos and the sorted coordinates of a few prepared sets."""
    import torch

    from syntropy_b200 import _engine as E
    from syntropy_b200 import workloads

    h = hashlib.sha1()
    rng = np.random.default_rng(9)
    sets = [workloads.c2_mi_1m(n=300_000)[0], workloads.c3_cmi(n=60_000)[0], rng.standard_normal((3, 5_000)),
            np.round(rng.standard_normal((2, 80_000)) * 8) / 8 + 0.0]   # (+ 0.0: no negative zeros, whose order is free)
    for rows in sets:
        pts = E.to_device_points(rows)
        prep = E.PreparedSet(pts, -1)
        torch.cuda.synchronize()
        h.update(prep.perm.cpu().numpy().tobytes())
        h.update(prep.sorted_f64.cpu().numpy().tobytes())
        dim, n, ld = prep.dim, prep.n, prep.ld
        vals = prep._view(prep.c.axis_vals, dim * ld * 8, torch.float64).view(dim, ld)[:, :n]
        pos = prep._view(prep.c.axis_pos, dim * ld * 4, torch.int32).view(dim, ld)[:, :n]
        h.update(vals.cpu().numpy().tobytes())
        h.update(pos.cpu().numpy().tobytes())
    return h.hexdigest()


def test_prepared_sets_identical_under_both_sorts():
    mine = _prepared_digest()
    code = "import sys; sys.path.insert(0, %r); from tests.test_gpu_sort import _prepared_digest; print(_prepared_digest())" % ROOT
    env = dict(os.environ, KSG_B200_SORT="sample")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == mine
