from __future__ import annotations

import numpy as np


def check_idxs(idxs, data):
    """None -> every channel (reference: syntropy/utils/utils.py:9-35)."""
    if idxs is None:
        return tuple(i for i in range(data.shape[0]))
    return tuple(idxs)


def is_chebyshev(p) -> bool:
    """p = inf: the maximum norm, the path BASELINE.json names (fp32-filtered engines).  Any other
    p >= 1 runs on the exact float64 Minkowski kernels (``syntropy_b200._minkowski``); p < 1 raises
    the ValueError cKDTree raises."""
    return p == np.inf
