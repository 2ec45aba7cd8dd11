"""Drop-in for the knn branch of ``syntropy.mixed.shannon`` (reference: syntropy/mixed/shannon.py).

Same signatures, return tuples (``ptw`` of shape (1, N), ``avg``) and assertions.  The reference
calls ``knn.differential_entropy(continuous_vars[:, mask], k)`` once per discrete class
(:82, :162, :241, :266); here the continuous array goes to the GPU once and every class -- plus the
pooled set where the estimator needs H(C) -- runs the KSG radius pipeline on a device-side gather of
its columns (``_engine.kl_entropy_groups``).  The class bookkeeping (np.unique, probabilities,
``-log p`` terms) stays on the host in the reference's float64 operation order.

``continuous_estimator="gaussian"`` -- the reference's DEFAULT -- is ``syntropy.gaussian``'s closed
form, not part of the KSG path and with nothing to accelerate: such calls are handed to the installed
reference (``syntropy.mixed``) unchanged, so default-argument calls keep working where syntropy is
installed; without it they raise ``NotImplementedError`` that says so.  The knn branch never falls
back to anything.
"""
from __future__ import annotations

import numpy as np

from .. import _engine as E

_ESTIMATORS = {"gaussian", "knn"}


def _gaussian_branch(name, *args, **kwargs):
    """The Gaussian plug-in is not on the KSG path: delegate to the reference if it is installed."""
    try:
        import importlib

        ref = importlib.import_module("syntropy.mixed")
    except Exception as exc:
        raise NotImplementedError(
            "syntropy_b200.mixed accelerates continuous_estimator='knn' (the KSG path); "
            "continuous_estimator='gaussian' (the default) is syntropy.gaussian's closed form and is "
            "delegated to the reference package, which is not importable here (%s). "
            "Pass continuous_estimator='knn' or install syntropy." % exc
        ) from exc
    return getattr(ref, name)(*args, **kwargs)


def _require_knn(continuous_estimator):
    assert continuous_estimator in _ESTIMATORS, "Continuous estimator options are 'gaussian' or 'knn'."
    return continuous_estimator == "knn"


_DENSE_KEY_LIMIT = 1 << 20  # integer class keys up to this many distinct codes are counted, not sorted


def _classes(discrete_vars):
    """``(probs, class_id, order, bounds)``: the distinct columns of ``discrete_vars`` in
    ``np.unique(axis=1)`` order (lexicographic, first row most significant; reference:
    mixed/shannon.py:53-56) with their probabilities; every sample's class; the sample indices
    class after class (ascending inside a class: what the reference's boolean masks select, :63-65)
    and the class boundaries in that list.

    ``np.unique(axis=1)`` sorts n structured records (0.4-0.9 s at n = 1e6): integer variables of
    small range are encoded as one mixed-radix key per sample and counted with ``bincount`` instead
    (same classes, same order); the grouping is a stable radix argsort of the 8/16-bit class ids."""
    dv = np.asarray(discrete_vars)
    n = dv.shape[1]
    class_id = counts = None
    if n and dv.dtype.kind in "iub" and dv.dtype != np.uint64:
        lo = [int(v) for v in dv.min(axis=1)]
        hi = [int(v) for v in dv.max(axis=1)]
        total = 1
        for a, b in zip(lo, hi):
            total *= b - a + 1
            if total > _DENSE_KEY_LIMIT:
                break
        if total <= _DENSE_KEY_LIMIT:
            key = np.zeros(n, dtype=np.int64)
            for r in range(dv.shape[0]):
                key *= hi[r] - lo[r] + 1
                key += dv[r].astype(np.int64) - lo[r]
            per_code = np.bincount(key, minlength=total)
            present = np.flatnonzero(per_code)
            counts = per_code[present]
            lut = np.zeros(total, dtype=np.int64)
            lut[present] = np.arange(present.shape[0])
            class_id = lut[key]
    if class_id is None:
        _, inverse, counts = np.unique(dv, axis=1, return_inverse=True, return_counts=True)
        class_id = np.asarray(inverse).reshape(-1)
    n_classes = int(counts.shape[0])
    narrow = np.uint8 if n_classes <= 256 else np.uint16 if n_classes <= 65536 else np.int64
    order = np.argsort(class_id.astype(narrow), kind="stable").astype(np.int64, copy=False)
    bounds = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    probs = counts / counts.sum()
    return probs, class_id, order, bounds


def _check_shapes(discrete_vars, continuous_vars):
    discrete_vars, continuous_vars = np.asarray(discrete_vars), np.asarray(continuous_vars)
    assert discrete_vars.shape[1] == continuous_vars.shape[1], (
        "The discrete and continuous variables must have the same number of samples."
    )
    return discrete_vars, continuous_vars


def _class_entropies(continuous_vars, order, bounds, k, pooled=False):
    """``knn.differential_entropy`` per class: the pointwise values in sample order, the pooled
    sample's pointwise values (or None), the per-class averages (``ptw.mean()`` = sum / n_class)
    and the pooled average."""
    by_sample, pooled_ptw, sums = E.kl_entropy_classes(continuous_vars, order, bounds, k, pooled)
    sizes = np.diff(bounds)
    with np.errstate(all="ignore"):
        avgs = [np.float64(sums[g]) / sizes[g] for g in range(sizes.shape[0])]
        pooled_avg = np.float64(sums[-1]) / continuous_vars.shape[1] if pooled else None
    return by_sample, pooled_ptw, avgs, pooled_avg


def _joint(probs, class_id, by_sample, avgs):
    """mixed/shannon.py:58-88: ptw = (0 - log p_class) + h_class(sample); avg = H(D) + sum_i p_i avg_i."""
    avg = -np.sum(probs * np.log(probs))
    with np.errstate(all="ignore"):
        ptw = ((0.0 - np.log(probs))[class_id] + by_sample).reshape(1, -1)
        conditional = 0.0
        for i in range(probs.shape[0]):
            conditional += probs[i] * avgs[i]
        avg += conditional
    return ptw, avg


def shannon_entropy(discrete_vars, continuous_vars, continuous_estimator="gaussian", k=5):
    """H(X, Y) = H(X) + H(Y | X) for discrete X, continuous Y (reference: mixed/shannon.py:7-90)."""
    discrete_vars, continuous_vars = _check_shapes(discrete_vars, continuous_vars)
    if not _require_knn(continuous_estimator):
        return _gaussian_branch("shannon_entropy", discrete_vars, continuous_vars, continuous_estimator=continuous_estimator, k=k)
    probs, class_id, order, bounds = _classes(discrete_vars)
    by_sample, _, avgs, _ = _class_entropies(continuous_vars, order, bounds, k)
    return _joint(probs, class_id, by_sample, avgs)


def conditional_entropy(discrete_vars, continuous_vars, conditional="discrete", continuous_estimator="gaussian", k=5):
    """H(X | Y) = H(X, Y) - H(Y) with either variable conditioning (reference: mixed/shannon.py:93-181)."""
    discrete_vars, continuous_vars = _check_shapes(discrete_vars, continuous_vars)
    assert conditional in {"discrete", "continuous"}, (
        "Please specify whether the conditioning variable is discrete or continuous."
    )
    if not _require_knn(continuous_estimator):
        return _gaussian_branch("conditional_entropy", discrete_vars, continuous_vars, conditional=conditional,
                                continuous_estimator=continuous_estimator, k=k)
    probs, class_id, order, bounds = _classes(discrete_vars)
    pooled = conditional == "continuous"
    by_sample, pooled_ptw, avgs, pooled_avg = _class_entropies(continuous_vars, order, bounds, k, pooled=pooled)
    ptw_joint, avg_joint = _joint(probs, class_id, by_sample, avgs)
    if pooled:
        ptw_marginal, avg_marginal = pooled_ptw.reshape(1, -1), pooled_avg
    else:
        ptw_marginal = (-np.log(probs))[class_id].reshape(1, -1)
        avg_marginal = -np.sum(probs * np.log(probs))
    with np.errstate(all="ignore"):
        return ptw_joint - ptw_marginal, avg_joint - avg_marginal


def mutual_information(discrete_vars, continuous_vars, continuous_estimator="gaussian", k=5):
    """I(D; C) = H(C) - H(C | D) (reference: mixed/shannon.py:184-275)."""
    discrete_vars, continuous_vars = _check_shapes(discrete_vars, continuous_vars)
    if not _require_knn(continuous_estimator):
        return _gaussian_branch("mutual_information", discrete_vars, continuous_vars, continuous_estimator=continuous_estimator, k=k)
    probs, class_id, order, bounds = _classes(discrete_vars)
    by_sample, pooled_ptw, avgs, avg = _class_entropies(continuous_vars, order, bounds, k, pooled=True)
    with np.errstate(all="ignore"):
        ptw = (pooled_ptw - by_sample).reshape(1, -1)
        conditional = 0.0
        for i in range(probs.shape[0]):
            conditional += probs[i] * avgs[i]
        avg = avg - conditional
    return ptw, avg
