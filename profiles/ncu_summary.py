"""Condense an .ncu-rep into the handful of metrics the roofline discussion needs.
usage: python profiles/ncu_summary.py <report.ncu-rep> [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "issue_active", "inst_executed_pipe_fp64", "pipe_fp64_cycles",
        "inst_executed_pipe_alu", "inst_executed_pipe_fma", "pipe_fmaheavy", "pipe_fmalite", "inst_executed_pipe_lsu",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp", "smsp__warps_issue_stalled", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "shared_mem", "launch__waves")


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        lines.append(f"== {name}  grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}")
        for h, u, v in zip(hdr, units, row):
            if any(k in h for k in KEYS) and v not in ("", "n/a"):
                lines.append(f"  {h} [{u}] = {v}")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
