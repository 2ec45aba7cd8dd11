#!/usr/bin/env python
"""Quantised inputs (ADC-style: multiples of a quantum, exact distance ties everywhere): time of one
KSG MI call and how many queries the fp32 filter handed to the float64 fallback."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntropy_b200 import _engine as E  # noqa: E402
from syntropy_b200 import knn  # noqa: E402

rng = np.random.default_rng(0)
for n, bits in ((200_000, 12), (1_000_000, 12), (1_000_000, 16), (1_000_000, 8)):
    x = rng.standard_normal(n)
    y = 0.6 * x + 0.8 * rng.standard_normal(n)
    q = 8.0 / (1 << bits)
    data = np.vstack([np.round(x / q) * q, np.round(y / q) * q])
    knn.mutual_information((0,), (1,), 5, data)  # warm-up at the full size (allocations, graph captures)
    knn.mutual_information((0,), (1,), 5, data)
    for key in E.stats:
        E.stats[key] = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ptw, avg = knn.mutual_information((0,), (1,), 5, data)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"n={n} {bits}-bit: {1e3 * dt:.1f} ms avg={avg!r} {dict(E.stats)}", flush=True)
