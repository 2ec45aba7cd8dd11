"""CPU checks of two small pieces of arithmetic the exactness of the fast path rests on.

* The per-query fp32 error band (csrc/fast_knn.cuh: band_query_f64): for the filter's arithmetic --
  rows centred in float64, rounded to fp32, subtracted in fp32 -- ``|d32 - d64| <= band(Mq, d64)``
  must hold for EVERY pair, whatever the scale, the offset of the data from the origin and the
  weight of its tails (numpy, all pairs of a few thousand points per regime).
* walk_tile (csrc/tile_queue.cuh), the near-first tile order of the planning pass: a permutation of
  the tiles with non-decreasing distance from the query block's own tile (compiled for the host).
"""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def band_constants():
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_knn.cuh")).read()
    m = re.search(r"return\s+([0-9.e+-]+)\s*\*\s*\(2\.0 \* mq \+ 2\.0 \* d\)\s*\*\s*([0-9.]+)\s*\+\s*([0-9.e+-]+);", src)
    assert m, "band_query_f64 changed shape: update this test with it"
    return float(m.group(1)), float(m.group(2)), float(m.group(3))


@pytest.mark.parametrize("regime", ["normal", "offset 1e6", "scale 1e-9", "cauchy", "lognormal sigma 3", "quantised"])
def test_fp32_error_band_holds_for_every_pair(regime):
    ulp, slack, floor = band_constants()
    assert ulp == 2.0 ** -24
    rng = np.random.default_rng(sum(map(ord, regime)))
    n, dim = 1500, 3
    x = rng.standard_normal((dim, n))
    if regime == "offset 1e6":
        x += 1e6 * np.arange(1, dim + 1)[:, None]
    elif regime == "scale 1e-9":
        x *= 1e-9
    elif regime == "cauchy":
        x = rng.standard_cauchy((dim, n))
    elif regime == "lognormal sigma 3":
        x = rng.lognormal(0.0, 3.0, (dim, n))
    elif regime == "quantised":
        x = np.round(x * 16) / 16 + 37.0
    center = np.sort(x, axis=1)[:, n // 2][:, None]           # per-coordinate median, as ksg_prepare
    shadow = (x - center).astype(np.float32)                   # __double2float_rn(__dsub_rn(v, centre))
    mq = np.abs(shadow).max(axis=0).astype(np.float64) * (1.0 + 1e-6) + 1e-300   # query_maxabs
    worst = 0.0
    for q in range(0, n, 7):
        d32 = np.abs(shadow[:, [q]] - shadow).max(axis=0).astype(np.float64)      # fp32 subtract, |.|, max
        d64 = np.abs(x[:, [q]] - x).max(axis=0)
        band = ulp * (2.0 * mq[q] + 2.0 * d64) * slack + floor
        assert np.all(np.abs(d32 - d64) <= band), regime
        worst = max(worst, float(np.max(np.abs(d32 - d64) / band)))
    assert worst <= 1.0
    if regime != "quantised":  # (multiples of 1/16 around 37 are exact in fp32: no error at all there)
        assert worst > 0.05    # the bound is not vacuous: real pairs come within a factor of 20 of it


def test_walk_tile_is_a_near_first_permutation(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "tile_queue.cuh")).read()
    a = src.index("__device__ __forceinline__ int64_t walk_tile")
    fn = src[a:src.index("\n}\n", a) + 3]
    (tmp_path / "walk.cpp").write_text("""
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define __device__
#define __forceinline__ inline
""" + fn + """
int main() {
    long bad = 0;
    for (int64_t n = 1; n <= 70; ++n)
        for (int64_t own = 0; own < n; ++own) {
            std::vector<int> seen(n, 0);
            int64_t prev = 0;
            for (int64_t p = 0; p < n; ++p) {
                const int64_t t = walk_tile(p, own, n);
                if (t < 0 || t >= n || seen[t]++) { ++bad; continue; }
                const int64_t dist = t > own ? t - own : own - t;
                if (dist < prev) ++bad;
                prev = dist;
            }
        }
    printf("%ld\\n", bad);
    return bad ? 1 : 0;
}
""")
    proc = subprocess.run([gxx, "-O1", "-std=c++17", "-o", "walk", "walk.cpp"], cwd=tmp_path, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    run = subprocess.run([str(tmp_path / "walk")], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.strip() == "0", run.stdout
