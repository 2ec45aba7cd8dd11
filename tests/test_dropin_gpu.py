"""Drop-in proof: user code written against ``syntropy.knn`` runs unchanged on
``syntropy_b200.knn`` installed under that name.

1. ``test_reference_assertions_through_the_syntropy_name``: every assertion of the reference's
   ``tests/test_knn.py`` (:29-113, incl. ``test_dkl`` at 2 x 1e6 samples) and
   ``tests/test_knn_temporal.py`` (:20-146, all ten Antropy constants) restated, reached through
   ``import syntropy.knn as knn`` / ``from syntropy.knn.temporal import ...``.
2. ``test_unmodified_reference_tests_pass_on_the_shim``: the reference's own test FILES, unmodified,
   executed by pytest with the shim as ``syntropy.knn`` -- when ``tools/stage_reference_tests.sh``
   has staged them under the git-ignored ``baseline/_ref/`` (the reference's sources are never part
   of this repository); skipped otherwise.
"""
import importlib
import os
import subprocess
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
ABS = 1e-6


def install_shim(reference_package_dir=None):
    """``syntropy`` -> a stub parent whose ``knn`` (and ``knn.shannon`` / ``multivariate_mi`` /
    ``temporal``) is the B200 implementation; with ``reference_package_dir`` every OTHER submodule
    (``syntropy.gaussian`` ...) resolves to the staged reference."""
    import syntropy_b200.knn as shim

    for name in [m for m in sys.modules if m == "syntropy" or m.startswith("syntropy.")]:
        del sys.modules[name]
    pkg = types.ModuleType("syntropy")
    pkg.__path__ = [reference_package_dir] if reference_package_dir else []
    sys.modules["syntropy"] = pkg
    sys.modules["syntropy.knn"] = shim
    pkg.knn = shim
    for sub in ("shannon", "multivariate_mi", "temporal"):
        sys.modules["syntropy.knn." + sub] = importlib.import_module("syntropy_b200.knn." + sub)
    return shim


def gaussian_kl(cov_posterior, cov_prior):
    """KL(N(0, cov_posterior) || N(0, cov_prior)) -- the closed form the reference's test_dkl compares
    with (syntropy/gaussian/shannon.py:381-440)."""
    k = cov_posterior.shape[0]
    inv = np.linalg.inv(cov_prior)
    return 0.5 * (np.trace(inv @ cov_posterior) - k + np.log(np.linalg.det(cov_prior) / np.linalg.det(cov_posterior)))


@pytest.fixture()
def syntropy_name():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    saved = {m: sys.modules[m] for m in list(sys.modules) if m == "syntropy" or m.startswith("syntropy.")}
    install_shim()
    yield
    for name in [m for m in sys.modules if m == "syntropy" or m.startswith("syntropy.")]:
        del sys.modules[name]
    sys.modules.update(saved)


def test_reference_assertions_through_the_syntropy_name(syntropy_name):
    import syntropy.knn as knn
    from scipy.special import erf
    from syntropy.knn.temporal import cross_sample_entropy, sample_entropy

    # ---- tests/test_knn.py:17-24 fixture, :29-75 known answers
    x = np.arange(100).reshape((1, 100))
    y = x[::-1]
    z = np.mod(x, 2)
    data = np.vstack((x, y, z))
    assert knn.differential_entropy(data=data, idxs=(0,), k=1)[1] == pytest.approx(5.870524698081722, ABS)
    assert knn.differential_entropy(data=data, idxs=(0,), k=4)[1] == pytest.approx(4.752310791249405, ABS)
    assert knn.mutual_information(idxs_x=(0,), idxs_y=(1,), k=1, data=data, algorithm=1)[1] == pytest.approx(5.177377517639623, ABS)
    assert knn.mutual_information(idxs_x=(0,), idxs_y=(1,), k=1, data=data, algorithm=2)[1] == pytest.approx(2.2173775176396227, ABS)
    assert knn.conditional_mutual_information(idxs_x=(0,), idxs_y=(1,), idxs_z=(2,), k=1, data=data)[1] == pytest.approx(4.479205338329424, ABS)
    assert knn.total_correlation(data=data, k=1)[-1] == pytest.approx(5.875549696949821, ABS)
    assert knn.dual_total_correlation(data=data, k=1)[-1] == pytest.approx(5.1773775176396235, ABS)
    assert knn.o_information(data=data, k=1)[-1] == pytest.approx(0.6981721793101983, ABS)
    assert knn.s_information(data=data, k=1)[-1] == pytest.approx(11.052927214589442, ABS)

    # ---- tests/test_knn.py:78-113 test_dkl: two 3-D Gaussian clouds of 1e6 samples, k = 1
    rng = np.random.default_rng(0)
    cov_prior = np.array([[0.99999999, 0.24404644, 0.65847509], [0.24404644, 0.99999985, 0.24163274],
                          [0.65847509, 0.24163274, 0.99999996]])
    cov_posterior = np.array([[0.99999997, 0.41352921, 0.3885488], [0.41352921, 1.00000004, 0.20463242],
                              [0.3885488, 0.20463242, 0.99999993]])
    prior_data = rng.multivariate_normal(mean=[0, 0, 0], cov=cov_prior, size=1_000_000).T
    posterior_data = rng.multivariate_normal(mean=[0, 0, 0], cov=cov_posterior, size=1_000_000).T
    ptw, avg = knn.kullback_leibler_divergence(posterior_data, prior_data, k=1)
    assert ptw.shape == (1_000_000,)
    assert avg == pytest.approx(gaussian_kl(cov_posterior, cov_prior), 10e-2)
    ptw, avg = knn.kullback_leibler_divergence(prior_data, posterior_data, k=1)
    assert avg == pytest.approx(gaussian_kl(cov_prior, cov_posterior), 10e-2)

    # ---- tests/test_knn_temporal.py:20-53 closed forms for i.i.d. processes (N = 50 000)
    N, tol = 50_000, 2e-3
    for m in (2, 3):
        xs = np.random.default_rng(0).standard_normal(N)
        assert sample_entropy((0,), xs[None, :], m=m, r=0.2, normalize_r=True) == pytest.approx(-np.log(erf(0.2 / 2)), abs=tol)
    xu = np.random.default_rng(0).random(N)
    r_abs = 0.2 * np.std(xu)
    assert sample_entropy((0,), xu[None, :], m=2, r=0.2, normalize_r=True) == pytest.approx(-np.log(2 * r_abs - r_abs ** 2), abs=tol)

    # ---- :55-76 the ten Antropy constants
    noise = np.random.default_rng(0).standard_normal((5, 10_000))
    want = [2.1960718154968815, 2.183291739784684, 2.183114546350932, 2.1764155747815264, 2.1901149239999116]
    for i in range(5):
        assert sample_entropy((i,), noise) == pytest.approx(want[i], abs=ABS)
    brownian = np.cumsum(noise, axis=-1)
    want = [0.04959120707043105, 0.061431721321519785, 0.07189084355966345, 0.08464506385060633, 0.0437523057152039]
    for i in range(5):
        assert sample_entropy((i,), brownian) == pytest.approx(want[i], abs=ABS)

    # ---- :79-146 cross sample entropy: brute force, symmetry, self-match offset, ordering
    from numpy.lib.stride_tricks import sliding_window_view
    from scipy.stats import zscore

    d2 = np.random.default_rng(0).standard_normal((2, 250))
    m, r = 2, 0.2
    est = cross_sample_entropy((0,), (1,), d2, m=m, r=r)
    n = d2.shape[1] - m
    sx, sy = zscore(d2[0]), zscore(d2[1])
    exs, eys = sliding_window_view(sx, m)[:n], sliding_window_view(sy, m)[:n]
    exb, eyb = sliding_window_view(sx, m + 1), sliding_window_view(sy, m + 1)
    count = lambda EX, EY: int((np.abs(EX[:, None, :] - EY[None, :, :]).max(axis=2) <= r).sum())  # noqa: E731
    assert est == pytest.approx(-np.log(count(exb, eyb) / count(exs, eys)), abs=ABS)
    d5 = np.random.default_rng(0).standard_normal((2, 5_000))
    assert cross_sample_entropy((0,), (1,), d5) == pytest.approx(cross_sample_entropy((1,), (0,), d5), abs=1e-12)
    x1 = np.random.default_rng(0).standard_normal((1, N))
    cross_xx, sampen = cross_sample_entropy((0,), (0,), x1), sample_entropy((0,), x1)
    assert cross_xx < sampen and cross_xx == pytest.approx(sampen, abs=2e-2)
    rng = np.random.default_rng(0)
    s1, s2, wn = np.cumsum(rng.standard_normal(20_000)), np.cumsum(rng.standard_normal(20_000)), rng.standard_normal(20_000)
    assert cross_sample_entropy((0,), (1,), np.vstack([s1, s2])) < cross_sample_entropy((0,), (1,), np.vstack([s1, wn]))


_RUNNER = r"""
import sys, types, importlib
root, staged = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
from tests.test_dropin_gpu import install_shim
install_shim(staged + "/syntropy")
import pytest
sys.exit(pytest.main(["-q", "-x", "-p", "no:cacheprovider", staged + "/ref_tests/test_knn.py",
                      staged + "/ref_tests/test_knn_temporal.py"]))
"""


def test_unmodified_reference_tests_pass_on_the_shim():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    if not os.path.exists(os.path.join(STAGED, "ref_tests", "test_knn.py")):
        pytest.skip("reference tests not staged (tools/stage_reference_tests.sh; needs the reference checkout)")
    out = subprocess.run([sys.executable, "-c", _RUNNER, ROOT, STAGED], capture_output=True, text=True, timeout=1800,
                         cwd=os.path.join(STAGED, "ref_tests"))
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-2000:]
