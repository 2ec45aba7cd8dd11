"""Synthetic inputs for the five BASELINE.json configurations.

All generators use ``np.random.default_rng(seed)`` and float64, with the draw order
written in SURVEY.md 8(d), so the CPU oracle, the golden fixtures, the GPU tests and
``bench.py`` see bit-identical data for the same (name, n, seed).

Each generator returns ``(data, call)`` where ``data`` is ``(n_variables, n_samples)``
and ``call`` describes the syntropy.knn call the configuration makes on it.
"""
from __future__ import annotations

import numpy as np


def _ar1(e, phi):
    """y[t] = phi*y[t-1] + e[t] along axis 1."""
    from scipy.signal import lfilter  # generator only; not on the product path

    return lfilter([1.0], [1.0, -phi], e, axis=1)


def c1_readme_mi(n=5_000, seed=0):
    """C1/C2: x ~ N(0,1), y = x^2 + 0.5*noise; KSG-1 MI 1D;1D, k=5 (README.md:95-106)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n)
    y = x ** 2 + 0.5 * rng.standard_normal(n)
    data = np.vstack([x, y])
    return data, {"fn": "mutual_information", "idxs_x": (0,), "idxs_y": (1,), "k": 5}


def c2_mi_1m(n=1_000_000, seed=0):
    """C2: the C1 generator at n = 1e6 (the configuration the metric is quoted on)."""
    return c1_readme_mi(n=n, seed=seed)


def c3_cmi(n=1_000_000, seed=0):
    """C3: Frenzel-Pompe CMI 2D;2D|2D, k=4."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((2, n))
    x = z + 0.5 * rng.standard_normal((2, n)) ** 3
    y = np.tanh(z) + 0.5 * x[::-1] + 0.5 * rng.laplace(size=(2, n))
    data = np.vstack([x, y, z])
    return data, {"fn": "conditional_mutual_information", "idxs_x": (0, 1),
                  "idxs_y": (2, 3), "idxs_z": (4, 5), "k": 4}


def c4_oinfo(n=250_000, seed=0):
    """C4: 8 variables sharing one latent factor; O-information / TC / DTC / S, k=4."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((1, n))
    data = 0.6 * s + rng.laplace(size=(8, n))
    return data, {"fn": "o_information", "k": 4}


def c5_series(n=4_000_000, seed=0):
    """The 4-channel coupled AR series behind C5, length n+1 (4 x (n+1))."""
    rng = np.random.default_rng(seed)
    e = rng.standard_normal((4, n + 1))
    a = _ar1(e, 0.7)
    a[1, 1:] += 0.5 * a[0, :-1]
    a[2, 1:] += 0.3 * np.tanh(a[1, :-1])
    a[3] = np.sign(a[3]) * np.abs(a[3]) ** 1.5
    return a


def c5_info_rate(n=4_000_000, seed=0):
    """C5: present || past (lag 1) embedding of the 4-channel series (8 x n);
    KSG-1 MI(present; past), k=5 -- the predictive-information rate proxy."""
    a = c5_series(n=n, seed=seed)
    data = np.vstack([a[:, 1:], a[:, :-1]])
    return data, {"fn": "mutual_information", "idxs_x": (0, 1, 2, 3),
                  "idxs_y": (4, 5, 6, 7), "k": 5}


def tie_fixture():
    """The reference's own integer, tie-saturated fixture (tests/test_knn.py:20-24):
    x = 0..99, y = x (``x[::-1]`` is a no-op on a (1,100) array), z = x mod 2."""
    x = np.arange(100).reshape((1, 100))
    y = x[::-1]
    z = np.mod(x, 2)
    return np.vstack((x, y, z))


CONFIGS = {
    "c1": c1_readme_mi,
    "c2": c2_mi_1m,
    "c3": c3_cmi,
    "c4": c4_oinfo,
    "c5": c5_info_rate,
}


def mixed_cases():
    """Discrete x continuous fixtures for the ``syntropy.mixed`` knn branch:
    ``name -> (discrete (vars, n) int64, continuous (vars, n) float64, k)``.
    "shift": balanced binary class, 1-D Gaussian with a class-dependent mean (the generator style of
    the reference's tests/test_mixed.py); "grid": two discrete variables (3 x 2 states, unbalanced),
    2-D continuous with class-dependent mean and scale; "small": one class has fewer than k + 1
    samples (its radii are inf, as cKDTree pads missing neighbours)."""
    rng = np.random.default_rng(21)
    cases = {}
    n = 3000
    d = rng.integers(0, 2, size=(1, n))
    c = rng.standard_normal((1, n)) + 1.5 * d
    cases["shift"] = (d.astype(np.int64), c, 5)
    n = 2500
    d = np.vstack([rng.choice(3, size=n, p=[0.5, 0.3, 0.2]), rng.integers(0, 2, size=n)])
    c = rng.standard_normal((2, n)) * (1.0 + 0.5 * d[1]) + np.vstack([d[0], -d[0]])
    cases["grid"] = (d.astype(np.int64), c, 3)
    n = 400
    d = np.zeros((1, n), dtype=np.int64)
    d[0, :3] = 7
    d[0, 3:150] = 2
    c = rng.standard_normal((3, n)) + 0.1 * d
    cases["small"] = (d, c, 4)
    return cases
