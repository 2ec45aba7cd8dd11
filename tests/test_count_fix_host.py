"""CPU check of the exactness argument of the fp32 count filter (csrc/fast_count.cuh): the filter's
per-query threshold and count_fix_kernel -- cut out of the header and compiled for the host
(tests/host_shims/count_fix_harness.cpp) -- must turn brute-force fp32 counts into the exact float64
strict / closed counts for every subspace, on centred, off-centre (1e4), heavy-tailed (Cauchy),
quantised and tiny-scale (1e-12) data, or flag the query; nothing may come out wrong."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_count_fix_kernel():
    knn = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_knn.cuh")).read()
    b0 = knn.index("__host__ __device__ inline double band_query_f64")
    band = knn[b0:knn.index("}", b0) + 1]
    r0 = knn.index("__host__ __device__ inline bool fp32_range_ok")
    band += "\n" + knn[r0:knn.index("\n}\n", r0) + 3]
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_count.cuh")).read()
    q0 = src.index("// Mq of the query at `row` of the fp32 shadow")
    helpers = src[q0:src.index("// ---- \"m < T\" as 1.0f / 0.0f on the FMA pipe.")]
    c0 = src.index("constexpr float COUNT_MIN_T")
    min_t = src[c0:src.index("\n", c0) + 1]
    f0 = src.index("constexpr int FIX_CAP")
    fix = src[f0:src.index("// ---------------------------------------------------------------- 1-D marginals")]
    return "\n".join([band, helpers, min_t, fix]).replace("fmaxf(", "fmaxf_(").replace("fabsf(", "std::fabs(")


def test_count_fix_kernel_makes_the_fp32_filter_exact_on_the_host(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    (tmp_path / "count_fix_kernel.inc").write_text(cut_count_fix_kernel())
    shutil.copy(os.path.join(ROOT, "tests", "host_shims", "count_fix_harness.cpp"), tmp_path / "count_fix_harness.cpp")
    exe = tmp_path / "count_fix_harness"
    proc = subprocess.run([gxx, "-O1", "-std=c++17", "-o", str(exe), "count_fix_harness.cpp"], cwd=tmp_path,
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    run = subprocess.run([str(exe), "120"], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:]
    words = run.stdout.split()
    stats = dict(zip(words[0::2], map(int, words[1::2])))
    assert stats["wrong"] == 0
    assert stats["checked"] > 100_000
    assert stats["corrected"] > 0          # the filter did over-count somewhere and the fix repaired it
    assert stats["flagged"] < stats["checked"] // 20
