#!/usr/bin/env python
"""Device timeline of one C2 step (diagnosis only -- numbers taken under a profiler are never
reported as results): every kernel / copy with its start offset, duration and the idle gap before
it, from torch.profiler (CUPTI).  Shows where a step's time goes between the kernels.

    python tools/timeline.py [--n 1000000] [--e2e] > gpurun_out/timeline.txt
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from syntropy_b200 import _engine as E  # noqa: E402
from syntropy_b200 import knn, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--e2e", action="store_true", help="profile the host-array API call instead of the resident step")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--cprofile", type=int, default=0,
                    help="instead of the device timeline: host profile (cProfile) of this many steps")
    args = ap.parse_args()
    data, call = workloads.c2_mi_1m(n=args.n, seed=0)
    k = call["k"]
    rows = data[list(call["idxs_x"]) + list(call["idxs_y"]), :]
    pts = E.to_device_points(rows)
    n = pts.shape[1]
    pinned = torch.empty(data.shape, dtype=torch.float64, pin_memory=True)
    pinned.numpy()[...] = data
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=pts.device)

    from syntropy_b200.knn.shannon import _xy_masks, mi1_program

    masks = _xy_masks(call["idxs_x"], call["idxs_y"])

    def step():
        if args.e2e:
            return knn.mutual_information(call["idxs_x"], call["idxs_y"], k, pinned.numpy())
        status = E.Status(pts.device)
        pw = E.ksg1_device(pts, k, masks, mi1_program, status=status)
        full, total = E.finish_device(pw, n, status=status)
        return E.settle(pw, n, status, full, total)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if args.cprofile:
        import cProfile
        import pstats

        prof = cProfile.Profile()
        prof.enable()
        for _ in range(args.cprofile):
            step()
        torch.cuda.synchronize()
        prof.disable()
        st = pstats.Stats(prof, stream=sys.stdout)
        st.sort_stats("tottime").print_stats(40)
        return
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            flush.zero_()
            step()
        torch.cuda.synchronize()
    events = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            events.append((ev.time_range.start, ev.time_range.end, ev.name))
    events.sort()
    # last step only: from the last flush kernel on
    flush_idx = [i for i, e in enumerate(events) if "FillFunctor<float>" in e[2]]
    if flush_idx:
        events = events[flush_idx[-1]:]
    t0 = events[0][1]
    prev_end = t0
    busy = 0.0
    print(f"{'start_us':>10} {'dur_us':>9} {'gap_us':>8}  name")
    for s, e, name in events[1:]:
        gap = s - prev_end
        print(f"{s - t0:10.1f} {e - s:9.1f} {gap:8.1f}  {name[:110]}")
        busy += e - s
        prev_end = max(prev_end, e)
    print(f"# span {prev_end - t0:.1f} us, sum of durations {busy:.1f} us (overlapping streams count twice)")


if __name__ == "__main__":
    main()
