#!/usr/bin/env python
"""Heavy-tailed inputs (range 1e4 .. 1e6 times the bulk): time of one KSG MI call and how many
queries left the fp32-filtered path for the float64 fallback (should be none: the fp32 copy is
centred on the median and the error band is per query)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntropy_b200 import knn, _engine as E
rng = np.random.default_rng(0)
n = 300_000
for name, x in (("student_t(1.2)", rng.standard_t(1.2, n)), ("lognormal(0,3)", rng.lognormal(0, 3, n)), ("normal", rng.standard_normal(n))):
    y = 0.5 * x + rng.standard_t(1.2, n) if name != "normal" else 0.5 * x + rng.standard_normal(n)
    data = np.vstack([x, y])
    for k in E.stats: E.stats[k] = 0
    knn.mutual_information((0,), (1,), 5, data)
    for k in E.stats: E.stats[k] = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ptw, avg = knn.mutual_information((0,), (1,), 5, data)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, "max|x|", float(np.abs(data).max()), "ms", round(1e3 * dt, 2), "avg", avg, dict(E.stats), flush=True)
