"""Drop-in for ``syntropy.knn.shannon`` (reference: syntropy/knn/shannon.py).

Same signatures, argument meaning, return tuples and assertion behaviour; the work is
done by the sm_100a kernels behind ``include/ksg_b200.h``.  ``p = inf`` (the maximum norm, the
path BASELINE.json names) runs on the fp32-filtered engines; any other ``p >= 1`` runs on the
exact float64 Minkowski kernels and reproduces the reference's forwarding of ``p`` function by
function (ignored by the joint query of algorithm 1, knn/shannon.py:186 vs :192).
"""
from __future__ import annotations

import numpy as np

from .. import _engine as E
from .. import _minkowski as M
from ._common import check_idxs, is_chebyshev


def differential_entropy(data, k, idxs=None, p=np.inf, noise_level=0.0, seed=0):
    """Kozachenko-Leonenko entropy (reference: knn/shannon.py:11-87).

    Returns ``(ptw (1, N) float64, average)`` with
    ``ptw = -psi(k) + psi(N) + d*log(2*eps)``.
    """
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    rows = data[idxs_, :]
    if noise_level > 0:
        # the reference jitters its fancy-index copy in place (:78-80)
        rng = np.random.default_rng(seed)
        rows += rng.standard_normal(rows.shape) * noise_level
    ptw, avg = E.kl_entropy(rows, k) if is_chebyshev(p) else M.entropy(rows, k, p)
    return ptw.reshape(1, -1), avg


def mutual_information(idxs_x, idxs_y, k, data, algorithm=1, p=np.inf):
    """KSG mutual information, algorithm 1 or 2 (reference: knn/shannon.py:90-134)."""
    assert algorithm in {1, 2}, "Algorithm must be 1 or 2."
    if algorithm == 1:
        return mutual_information_1(idxs_x=idxs_x, idxs_y=idxs_y, k=k, data=data, p=p)
    return mutual_information_2(idxs_x=idxs_x, idxs_y=idxs_y, k=k, data=data, p=p)


def _xy_masks(idxs_x, idxs_y):
    dx, dy = len(idxs_x), len(idxs_y)
    return [E.mask_of(range(dx)), E.mask_of(range(dx, dx + dy))]


def mi1_program(psi_k, psi_n):
    """(psi_k + psi_N) - psi(n_x+1) - psi(n_y+1), in that order (knn/shannon.py:188-194)."""
    prog = E.PsiProgram(init=psi_k + psi_n)
    prog.add(0, -1)
    prog.add(1, -1)
    return prog


def mutual_information_1(idxs_x, idxs_y, k, data, p=np.inf):
    """KSG-1: ``(psi(k)+psi(N)) - psi(n_x+1) - psi(n_y+1)`` (reference: knn/shannon.py:137-196)."""
    data = np.asarray(data)
    idxs_x, idxs_y = tuple(idxs_x), tuple(idxs_y)
    if not is_chebyshev(p):
        # the joint radius stays a maximum-norm radius (:186), only the counts see p (:192)
        ptw, avg = M.ksg1(data[idxs_x + idxs_y, :], k, _xy_masks(idxs_x, idxs_y), mi1_program, p, joint_chebyshev=True)
        return ptw.reshape(1, -1), avg
    rows = E.select_rows(data, idxs_x + idxs_y)
    ptw, avg = E.ksg1_family(rows, k, _xy_masks(idxs_x, idxs_y), mi1_program)
    return ptw.reshape(1, -1), avg


def mutual_information_2(idxs_x, idxs_y, k, data, p=np.inf):
    """KSG-2: per-marginal radii from the k joint neighbours, closed counts,
    ``psi(k) - 1/k + psi(N) - psi(n_x) - psi(n_y)`` (reference: knn/shannon.py:199-265).
    Neighbour identity among exactly tied distances is cKDTree-internal in the reference;
    here ties are broken by index, so bit parity holds on tie-free data."""
    data = np.asarray(data)
    idxs_x, idxs_y = tuple(idxs_x), tuple(idxs_y)

    def program(psi_k, psi_n):
        prog = E.PsiProgram(init=psi_k - (1 / k) + psi_n)
        prog.add(0, -1, count_offset=0)
        prog.add(1, -1, count_offset=0)
        return prog

    if not is_chebyshev(p):
        ptw, avg = M.ksg2(data[idxs_x + idxs_y, :], k, _xy_masks(idxs_x, idxs_y), program, p)
        return ptw.reshape(1, -1), avg
    rows = E.select_rows(data, idxs_x + idxs_y)
    ptw, avg = E.ksg2_family(rows, k, _xy_masks(idxs_x, idxs_y), program)
    return ptw.reshape(1, -1), avg


def conditional_mutual_information(idxs_x, idxs_y, idxs_z, k, data, p=np.inf):
    """Frenzel-Pompe CMI: ``psi(k) + psi(n_z+1) - psi(n_xz+1) - psi(n_yz+1)``
    (reference: knn/shannon.py:268-347)."""
    data = np.asarray(data)
    idxs_x, idxs_y, idxs_z = tuple(idxs_x), tuple(idxs_y), tuple(idxs_z)
    masks, program = cmi_spec(len(idxs_x), len(idxs_y), len(idxs_z))
    if not is_chebyshev(p):
        ptw, avg = M.ksg1(data[idxs_x + idxs_y + idxs_z, :], k, masks, program, p, joint_chebyshev=False)  # :332,338
        return ptw.reshape(1, -1), avg
    rows = E.select_rows(data, idxs_x + idxs_y + idxs_z)
    ptw, avg = E.ksg1_family(rows, k, masks, program)
    return ptw.reshape(1, -1), avg


def cmi_spec(dx, dy, dz):
    """(subspace masks z | xz | yz over the x,y,z-concatenated rows, psi program) of the CMI."""
    x = list(range(dx))
    y = list(range(dx, dx + dy))
    z = list(range(dx + dy, dx + dy + dz))
    masks = [E.mask_of(z), E.mask_of(x + z), E.mask_of(y + z)]

    def program(psi_k, psi_n):
        prog = E.PsiProgram(init=psi_k)
        prog.add(0, +1)
        prog.add(1, -1)
        prog.add(2, -1)
        return prog

    return masks, program


def kullback_leibler_divergence(posterior_data, prior_data, k, p=np.inf):
    """kNN KL divergence ``d*log(nu/rho) + log(m/(n-1))`` (reference: knn/shannon.py:350-410).
    Returns ``(ptw (n,), average)``."""
    posterior_data = np.asarray(posterior_data)
    prior_data = np.asarray(prior_data)
    assert posterior_data.shape[0] == prior_data.shape[0], (
        "The data must have the same number of dimensions."
    )
    if not is_chebyshev(p):
        return M.kl_divergence(posterior_data, prior_data, k, p)
    ptw, avg = E.kl_divergence(posterior_data, prior_data, k)
    return ptw, avg
