#!/usr/bin/env python
"""One rank's share of the sharded C2 step on ONE GPU (no process group): prepared set of all n points,
k-NN / 1-D counts / psi for the query slice [0, n / G) only.  Under `ncu --metrics gpu__time_duration.sum`
this gives the per-kernel times of a rank of a G-GPU run, which ncu cannot capture from a multi-rank launch.

    python tools/slice_step.py --world 8 [--steps 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from syntropy_b200 import _engine as E  # noqa: E402
from syntropy_b200 import workloads  # noqa: E402
from syntropy_b200.knn.shannon import _xy_masks, mi1_program  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--n", type=int, default=1_000_000)
    args = ap.parse_args()
    data, call = workloads.c2_mi_1m(n=args.n)
    rows = data[call["idxs_x"] + call["idxs_y"], :]
    pts = E.to_device_points(rows)
    n, k = pts.shape[1], call["k"]
    masks = _xy_masks(call["idxs_x"], call["idxs_y"])
    c = -(-n // args.world)
    q0, q1 = min(n, args.rank * c), min(n, (args.rank + 1) * c)
    prog = mi1_program(E.psi_host(k), E.psi_host(n))
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=pts.device)
    E.timer.enabled = True
    for it in range(args.steps + 3):
        if it == 3:
            E.timer.reset()
        flush.zero_()
        status = E.Status(pts.device)
        pw = E._ksg1_step_ext(pts, k, masks, prog, n, q0, q1, True, None, status)
        torch.cuda.synchronize()
    tot = E.timer.totals_ms()
    print(f"rank {args.rank}/{args.world}", {kk: round(sum(v) / args.steps, 4) for kk, v in tot.items()}, "nflag", int(status.nflag.item()))


if __name__ == "__main__":
    main()
