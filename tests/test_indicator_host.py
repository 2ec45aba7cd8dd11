"""CPU check of the saturating-FMA indicator of the fused count kernel (csrc/fast_count.cuh:
make_ind_scale / ind_lt): exactly 1.0 for m < T and 0.0 otherwise over the whole exponent range,
one ulp either side of T, infinities and NaN (3 million cases, tests/host_shims/indicator_harness.cpp)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_indicator():
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_count.cuh")).read()
    a = src.index("constexpr float COUNT_MIN_T")
    text = src[a:src.index("// ---------------------------------------------------------------- subspace specs", a)]
    text, n = re.subn(r'asm\("fma\.rn\.sat\.f32.*?\)\);', "r = fma_sat(m, s.neg_b, s.tb);", text, flags=re.S)
    assert n == 1, "the indicator is expected to be one fma.rn.sat.f32"
    return text


def test_saturating_fma_indicator_is_exact(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    (tmp_path / "indicator.inc").write_text(cut_indicator())
    shutil.copy(os.path.join(ROOT, "tests", "host_shims", "indicator_harness.cpp"), tmp_path / "indicator_harness.cpp")
    exe = tmp_path / "indicator_harness"
    proc = subprocess.run([gxx, "-O1", "-std=c++17", "-o", str(exe), "indicator_harness.cpp"], cwd=tmp_path,
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-3000:]
    words = run.stdout.split()
    stats = dict(zip(words[0::2], map(int, words[1::2])))
    assert stats["wrong"] == 0 and stats["checked"] > 2_500_000
