"""Drop-in for ``syntropy.knn.multivariate_mi`` (reference: syntropy/knn/multivariate_mi.py).

Every function needs ONE joint radius vector and then the strict counts of several
marginal subspaces at that radius; all of them come from a single
``ksg_count_subspaces`` call instead of one kd-tree and one Python loop per marginal.
"""
from __future__ import annotations

import numpy as np

from .. import _engine as E
from ._common import check_idxs


def total_correlation(data, k, idxs=None, algorithm=1):
    """Total correlation, KSG algorithm 1 or 2 (reference: knn/multivariate_mi.py:11-51)."""
    assert algorithm in {1, 2}, "Algorithm must be 1 or 2."
    if algorithm == 1:
        return total_correlation_1(data=data, k=k, idxs=idxs)
    return total_correlation_2(data=data, k=k, idxs=idxs)


def total_correlation_1(data, k, idxs=None):
    """``psi(k) + (m-1)psi(N) - sum_d psi(n_d+1)`` (reference: knn/multivariate_mi.py:54-115)."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    rows = E.select_rows(data, idxs_)
    masks, program = spec("total_correlation", m)
    ptw, avg = E.ksg1_family(rows, k, masks, program)
    return ptw.reshape(1, -1), avg


def total_correlation_2(data, k, idxs=None):
    """KSG-2 total correlation (reference: knn/multivariate_mi.py:118-193)."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    rows = E.select_rows(data, idxs_)

    def program(psi_k, psi_n):
        prog = E.PsiProgram(init=psi_k - ((m - 1) / k) + ((m - 1) * psi_n))
        for d in range(m):
            prog.add(d, -1, count_offset=0)
        return prog

    ptw, avg = E.ksg2_family(rows, k, [1 << d for d in range(m)], program)
    return ptw.reshape(1, -1), avg


def _leave_one_out(m, d):
    return ((1 << m) - 1) & ~(1 << d)


def dual_total_correlation(data, k, idxs=None):
    """``(m-1)*[psi(k)-psi(N) - sum_i (psi(c_i+1)-psi(N))/(m-1)]`` with c_i counted in the
    leave-one-out subspaces (reference: knn/multivariate_mi.py:196-260)."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    rows = E.select_rows(data, idxs_)
    masks, program = spec("dual_total_correlation", m)
    with np.errstate(all="ignore"):
        ptw, avg = E.ksg1_family(rows, k, masks, program)
    return ptw.reshape(1, -1), avg


def _big_small_masks(m):
    masks = []
    for d in range(m):
        masks.append(_leave_one_out(m, d))  # 2d   : big marginal
        masks.append(1 << d)                # 2d+1 : small marginal
    return masks


def spec(name, m):
    """(subspace masks over the m selected rows, psi program) of a multivariate estimator; the
    programs replay the reference's float64 update order (knn/multivariate_mi.py:100-113,
    241-258, 311-333, 394-413)."""
    if name == "total_correlation":
        def program(psi_k, psi_n):
            prog = E.PsiProgram(init=psi_k + (m - 1) * psi_n)
            for d in range(m):
                prog.add(d, -1)
            return prog
        return [1 << d for d in range(m)], program
    if name == "dual_total_correlation":
        def program(psi_k, psi_n):
            prog = E.PsiProgram(init=psi_k - psi_n, psi_n=psi_n, scale=float(m - 1))
            for i in range(m):
                prog.add(i, -1, centered=True, divisor=m - 1)
            return prog
        return [_leave_one_out(m, i) for i in range(m)], program
    if name == "s_information":
        def program(psi_k, psi_n):
            prog = E.PsiProgram(init=psi_k - psi_n, psi_n=psi_n, scale=float(m))
            for d in range(m):
                prog.add(2 * d, -1, centered=True, divisor=m)
                prog.add(2 * d + 1, -1, centered=True, divisor=m)
            return prog
        return _big_small_masks(m), program
    if name == "o_information":
        def program(psi_k, psi_n):
            prog = E.PsiProgram(init=psi_k - psi_n, psi_n=psi_n, scale=float(2 - m))
            for d in range(m):
                prog.add(2 * d, -1, centered=True, divisor=m - 2)
                prog.add(2 * d + 1, +1, centered=True, divisor=m - 2)
            return prog
        return _big_small_masks(m), program
    raise ValueError(name)


def s_information(data, k, idxs=None):
    """S-information (reference: knn/multivariate_mi.py:263-335)."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    m = len(idxs_)
    # (the reference selects the small marginal by the SLICE ``idxs_[d]:idxs_[d]+1`` and the big one by
    #  fancy index, knn/multivariate_mi.py:317,321,398,402: the same rows for non-negative indices)
    rows = E.select_rows(data, idxs_)
    masks, program = spec("s_information", m)
    ptw, avg = E.ksg1_family(rows, k, masks, program)
    return ptw.reshape(1, -1), avg


def o_information(data, k, idxs=None):
    """O-information (reference: knn/multivariate_mi.py:338-415); for two variables the
    reference returns ``(np.zeros(N), 0.0)`` without touching the data (:385-386)."""
    data = np.asarray(data)
    idxs_ = check_idxs(idxs, data)
    N = data.shape[1]
    m = len(idxs_)
    if m == 2:
        return np.zeros(N), 0.0
    rows = E.select_rows(data, idxs_)
    masks, program = spec("o_information", m)
    with np.errstate(all="ignore"):
        ptw, avg = E.ksg1_family(rows, k, masks, program)
    return ptw.reshape(1, -1), avg
