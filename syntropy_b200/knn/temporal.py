"""Drop-in for ``syntropy.knn.temporal`` (reference: syntropy/knn/temporal.py)."""
from __future__ import annotations

import numpy as np

from .. import _engine as E
from .. import _minkowski as M
from ._common import is_chebyshev


def sample_entropy(idx, data, m=2, r=0.2, normalize_r=True, norm=np.inf):
    """Sample entropy ``-log(A/B)`` from closed-ball ordered pair counts of the (m+1)- and
    m-templates, self pairs removed (reference: knn/temporal.py:10-77)."""
    series = np.asarray(data)[idx[0], :]
    N = series.shape[0] - m
    if normalize_r:
        r = r * np.std(series)
    small, big = E.pair_counts(series, series, m, r) if is_chebyshev(norm) else M.pair_counts(series, series, m, r, norm)
    counts_small = small - N
    counts_big = big - N
    if counts_big == 0:
        return np.inf
    with np.errstate(all="ignore"):
        return -np.log(counts_big / counts_small)


def _zscore(x):
    # scipy.stats.zscore defaults: (x - mean) / std with ddof=0
    x = np.asarray(x)
    return (x - x.mean()) / x.std()


def cross_sample_entropy(idx_x, idx_y, data, m=2, r=0.2, norm=np.inf):
    """Cross sample entropy of two z-scored channels; no self-pair removal
    (reference: knn/temporal.py:79-148)."""
    data = np.asarray(data)
    series_x = _zscore(data[idx_x[0], :])
    series_y = _zscore(data[idx_y[0], :])
    B, A = E.pair_counts(series_x, series_y, m, r) if is_chebyshev(norm) else M.pair_counts(series_x, series_y, m, r, norm)
    if A == 0:
        return np.inf
    with np.errstate(all="ignore"):
        return -np.log(A / B)
