#!/usr/bin/env python
"""Per-CUDA-source-line summary of one kernel of an .ncu-rep captured with `--import-source on`:
warp instructions executed and stall samples, aggregated from `ncu --page source --print-source
cuda,sass --csv`.

    python tools/ncu_source_lines.py report.ncu-rep <kernel regex> [min_pct]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    per = {}
    fname = ""
    for row in rows:
        if row and row[0] == "File Path":
            fname = row[1].rsplit("/", 1)[-1] if len(row) > 1 else ""
            continue
        if row and row[0] == "Line No":
            # one section per source file that contributed lines to the kernel (inlined headers)
            hdr = row
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(row) < len(hdr) - 2 or not row[0].strip().isdigit():
            continue  # (rows without a line number are the SASS instructions under a CUDA line)
        key = (fname + ":" + row[0].strip(), row[1].strip())
        ent = per.setdefault(key, [0, 0, {}])
        try:
            ent[0] += int(row[i_inst] or 0)
            ent[1] += int(row[i_samp] or 0)
            for h, i in stall_cols:
                v = int(row[i] or 0)
                if v:
                    ent[2][h] = ent[2].get(h, 0) + v
        except ValueError:
            pass
    tot_i = sum(v[0] for v in per.values()) or 1
    tot_s = sum(v[1] for v in per.values()) or 1
    print(f"# total: {tot_i} warp instructions, {tot_s} samples")
    print("#  inst%  samp%  line  top stalls | source")
    for (line, src), (ni, ns, st) in sorted(per.items(), key=lambda kv: -kv[1][0] - kv[1][1] * tot_i / tot_s):
        pi, ps = 100.0 * ni / tot_i, 100.0 * ns / tot_s
        if pi < min_pct and ps < min_pct:
            continue
        top = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{pi:6.1f} {ps:6.1f}  {line:>24}  {top:40s} | {src[:110]}")


if __name__ == "__main__":
    main()
