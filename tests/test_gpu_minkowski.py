"""Minkowski p other than the maximum norm on the GPU, against
tests/golden/minkowski.npz -- outputs of the UNMODIFIED reference (make_golden.py --only-minkowski) --
and the CPU oracle.  p = 1, 2: radii and counts bit-exact, nats within 1e-12 relative.  Other p
(here 3): the device pow() is not the host libm's, so radii agree to a few ulps and a count may
differ by one where a pair sits within those ulps of the boundary."""
import os

import numpy as np
import pytest

from oracle import ksg_oracle as orc
from tests.test_gpu_parity import close_pair
from tests.test_oracle import _minkowski_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from syntropy_b200 import knn as mod

    return mod


@pytest.mark.parametrize("p", [1, 2])
def test_minkowski_exact_p(knn, golden, p):
    g = golden("minkowski")
    cont, tied, post, prior, series = _minkowski_inputs()
    t = f"p{p}"
    with np.errstate(all="ignore"):
        ptw, avg = knn.differential_entropy(cont, 4, idxs=(0, 2, 5), p=p)
        close_pair(ptw, g[f"{t}_entropy_ptw"], avg, g[f"{t}_entropy_avg"])
        for alg in (1, 2):
            ptw, avg = knn.mutual_information((0, 1), (2, 3), 4, cont, algorithm=alg, p=p)
            close_pair(ptw, g[f"{t}_mi{alg}_ptw"], avg, g[f"{t}_mi{alg}_avg"])
        ptw, avg = knn.mutual_information((0,), (1,), 3, cont, algorithm=1, p=p)
        close_pair(ptw, g[f"{t}_mi1_1d_ptw"], avg, g[f"{t}_mi1_1d_avg"])
        ptw, avg = knn.conditional_mutual_information((0, 1), (2,), (3, 4), 3, cont, p=p)
        close_pair(ptw, g[f"{t}_cmi_ptw"], avg, g[f"{t}_cmi_avg"])
        ptw, avg = knn.kullback_leibler_divergence(post, prior, k=3, p=p)
        close_pair(ptw, g[f"{t}_kl_ptw"], avg, g[f"{t}_kl_avg"])
        se = np.array([knn.sample_entropy((c,), series, m=2, r=0.3, norm=p) for c in range(2)])
        assert np.array_equal(se, g[f"{t}_sampen"])
        assert knn.cross_sample_entropy((0,), (1,), series, m=2, r=0.35, norm=p) == g[f"{t}_cross_sampen"]
        # tie-saturated data: psi(0) = -inf terms must land in the same places
        ptw, avg = knn.mutual_information((0, 1), (2,), 3, tied, algorithm=1, p=p)
        close_pair(ptw, g[f"{t}_tied_mi1_ptw"], avg, g[f"{t}_tied_mi1_avg"])
        ptw, avg = knn.conditional_mutual_information((0,), (1,), (2, 3), 2, tied, p=p)
        close_pair(ptw, g[f"{t}_tied_cmi_ptw"], avg, g[f"{t}_tied_cmi_avg"])


@pytest.mark.parametrize("p", [1, 2])
def test_minkowski_radii_and_counts_bit_exact(golden, p):
    from syntropy_b200 import _engine as E
    from syntropy_b200 import _minkowski as M

    g = golden("minkowski")
    cont = _minkowski_inputs()[0]
    pts = E.to_device_points(cont[(0, 1, 2), :])
    n = pts.shape[1]
    eps = M.knn_radius_p(pts, 5, 0, n, p)
    assert np.array_equal(eps.cpu().numpy(), g[f"p{p}_eps"])
    sub = E.to_device_points(cont[(0, 1), :])
    strict = M.count_subspaces_p(sub, [0b11], eps, 0, n, p, strict=True).cpu().numpy()[0]
    closed = M.count_subspaces_p(sub, [0b11], eps, 0, n, p, strict=False).cpu().numpy()[0]
    assert np.array_equal(strict, g[f"p{p}_counts_strict"])
    assert np.array_equal(closed, g[f"p{p}_counts_closed"])
    # higher-dimensional sums (the four-way partial sums of p = 2 start at 4 coordinates) and big k
    rng = np.random.default_rng(4)
    rows = rng.standard_normal((7, 1200)) + 0.3 * rng.standard_normal((1, 1200))
    pts = E.to_device_points(rows)
    eps, idx = M.knn_radius_p(pts, 9, 0, 1200, p, want_indices=True)
    ref_eps, ref_idx = orc.knn_radius_p(rows, 8, p, return_indices=True)
    assert np.array_equal(eps.cpu().numpy(), ref_eps)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    masks = [0b1111111, 0b0011111, 0b1100000, 0b0101010]
    got = M.count_subspaces_p(pts, masks, eps, 0, 1200, p).cpu().numpy()
    for row, m in enumerate(masks):
        dims = [d for d in range(7) if (m >> d) & 1]
        assert np.array_equal(got[row], orc.count_within_p(rows[dims, :], ref_eps, p)), (p, m)


def test_minkowski_general_p(knn, golden):
    """p = 3: device pow() against libm pow() -- radii to 1e-13 relative, nearly all counts equal."""
    from syntropy_b200 import _engine as E
    from syntropy_b200 import _minkowski as M

    g = golden("minkowski")
    cont, tied, post, prior, series = _minkowski_inputs()
    pts = E.to_device_points(cont[(0, 1, 2), :])
    n = pts.shape[1]
    eps = M.knn_radius_p(pts, 5, 0, n, 3.0).cpu().numpy()
    assert np.allclose(eps, g["p3_eps"], rtol=1e-13, atol=0)
    sub = E.to_device_points(cont[(0, 1), :])
    strict = M.count_subspaces_p(sub, [0b11], E.to_device_points(g["p3_eps"][None, :])[0], 0, n, 3.0).cpu().numpy()[0]
    assert np.mean(strict != g["p3_counts_strict"]) < 0.005 and np.abs(strict - g["p3_counts_strict"]).max() <= 1
    with np.errstate(all="ignore"):
        for name, call in {
            "entropy": lambda: knn.differential_entropy(cont, 4, idxs=(0, 2, 5), p=3),
            "mi1": lambda: knn.mutual_information((0, 1), (2, 3), 4, cont, algorithm=1, p=3),
            "cmi": lambda: knn.conditional_mutual_information((0, 1), (2,), (3, 4), 3, cont, p=3),
            "kl": lambda: knn.kullback_leibler_divergence(post, prior, k=3, p=3),
        }.items():
            ptw, avg = call()
            assert avg == pytest.approx(float(g[f"p3_{name}_avg"]), rel=2e-3), name
    with pytest.raises(ValueError):
        knn.differential_entropy(cont, 4, p=0.5)   # cKDTree: "Only p-norms with 1<=p<=infinity permitted"
