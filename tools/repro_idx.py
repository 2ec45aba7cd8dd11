#!/usr/bin/env python
"""Repro helper for the neighbour-identity path of the windowed k-NN stage (KSG-2): regenerate one case of
tools/fuzz_engines.py and compare ``fast_knn(..., want_indices=True)`` -- first pass and settled -- with the
exact float64 kernel, several times over, printing what differs (radius, neighbour set, first-pass flag,
work units of the query's block).

    python tools/repro_idx.py --seed 21 --case 170 [--repeat 8]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fuzz_engines as F  # noqa: E402
from syntropy_b200 import _engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=21)
    ap.add_argument("--case", type=int, default=170)
    ap.add_argument("--repeat", type=int, default=8)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    for case in range(args.case + 1):
        dim = int(rng.integers(1, 9))
        n = int(rng.choice([37, 200, 511, 513, 1000, 2049, 5000, 9973, 20011]))
        k = int(rng.integers(1, min(12, n - 1)))
        kind, rows = F.make_data(rng, dim, n)
        F.subspaces(rng, dim)
    print(f"case {args.case}: {kind} dim={dim} n={n} k={k}", flush=True)
    pts = E.to_device_points(rows)
    kp = k + 1
    for rep in range(args.repeat):
        prep = E.PreparedSet(pts, E.order_for(dim))
        srt = prep.sorted_f64.contiguous()
        eps_x, idx_x = E.knn_radius(srt, kp, 0, n, want_indices=True)   # exact, in the prepared order
        eps1, idx1, flags1, nflag1 = E._fast_knn_call(prep, kp, 0, n, True, want_indices=True)
        torch.cuda.synchronize()
        f1 = flags1.cpu().numpy()
        ok = f1 == 0
        e_bad = np.flatnonzero(ok & (eps1.cpu().numpy() != eps_x.cpu().numpy()))
        a, b = idx1.cpu().numpy(), idx_x.cpu().numpy()
        i_bad = np.flatnonzero(ok & (a != b).any(axis=1))
        eps2, idx2 = E.fast_knn(prep, kp, 0, n, True, want_indices=True)
        torch.cuda.synchronize()
        s_bad = np.flatnonzero((idx2.cpu().numpy() != b).any(axis=1) | (eps2.cpu().numpy() != eps_x.cpu().numpy()))
        print(f"rep {rep}: first pass flagged {int((f1 != 0).sum())} (nflag {int(nflag1.item())}), decided-but-wrong radii {e_bad.size}, "
              f"neighbour lists {i_bad.size}; settled wrong {s_bad.size}", flush=True)
        for q in list(i_bad[:4]) + [s for s in s_bad[:4] if s not in i_bad[:4]]:
            dx = (srt[:, a[q]] - srt[:, q:q + 1]).abs().max(dim=0).values.cpu().numpy() if a[q].min() >= 0 else None
            dr = (srt[:, b[q]] - srt[:, q:q + 1]).abs().max(dim=0).values.cpu().numpy()
            print(f"   q={q} block={q // 128} flag1={f1[q]} eps1={float(eps1[q])!r} eps_x={float(eps_x[q])!r}\n"
                  f"      got  idx {a[q].tolist()} dist {None if dx is None else dx.tolist()}\n"
                  f"      want idx {b[q].tolist()} dist {dr.tolist()}", flush=True)


if __name__ == "__main__":
    main()
