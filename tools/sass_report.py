#!/usr/bin/env python
"""SASS evidence for the claims of DESIGN.md: per kernel of libksg_b200.so, how many bulk-copy
(UBLKCP), mbarrier (SYNCS.*), packed-fp32 (FADD2 / FMUL2 / FFMA2), three-input min/max (FMNMX3),
fp64 (DADD / DMUL / DSETP) and tensor-core (UTCMMA / HMMA -- expected: none, a max-norm distance is
not a contraction) instructions the sm_100a code contains.  No GPU needed.

    python tools/sass_report.py > profiles/r2_sass_report.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "syntropy_b200", "lib", "libksg_b200.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS\."), ("FADD2", r"\bFADD2"),
            ("FMNMX3", r"\bFMNMX3"), ("FMNMX", r"\bFMNMX\b"), ("FSET", r"\bFSET[P.]"), ("SHF", r"\bSHF\."),
            ("DADD/DMUL", r"\bD(ADD|MUL|FMA)\b"), ("DSETP", r"\bDSETP"), ("ATOM/RED", r"\b(ATOM|ATOMG|RED)\."),
            ("BAR", r"\bBAR\."), ("LDL/STL", r"\b(LDL|STL)\b"), ("tensor", r"\b(UTCMMA|UTCHMMA|HMMA|IMMA|QGMMA)\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        per[cur]["instructions"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                per[cur][name] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} functions, architectures {arch}")
    cols = ["instructions"] + [n for n, _ in PATTERNS]
    print("# " + " ".join(f"{c:>9s}" for c in cols) + "  kernel")
    total = collections.Counter()
    rows = []
    for (fn, cnt), name in zip(per.items(), demangled):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("ksg::", "")
        total.update(cnt)
        rows.append((short, cnt))
    keep = [r for r in rows if any(k in r[0] for k in ("knn_", "count_", "axis_count", "psi_", "sum_stage", "paircount",
                                                       "curve_", "ss_", "gather_sorted", "tile_box", "refine"))]
    for short, cnt in sorted(keep, key=lambda r: r[0]):
        print("  " + " ".join(f"{cnt[c]:9d}" for c in cols) + "  " + short[:90])
    print("# all functions of the library together:")
    print("  " + " ".join(f"{total[c]:9d}" for c in cols))


if __name__ == "__main__":
    sys.exit(main())
