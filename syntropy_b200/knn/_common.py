from __future__ import annotations

import numpy as np


def check_idxs(idxs, data):
    """None -> every channel (reference: syntropy/utils/utils.py:9-35)."""
    if idxs is None:
        return tuple(i for i in range(data.shape[0]))
    return tuple(idxs)


def require_chebyshev(p):
    if not (p == np.inf):
        raise NotImplementedError(
            "syntropy_b200 implements the Chebyshev (p=inf) KSG path only; "
            f"p={p!r} is not supported and there is no CPU fallback"
        )
