#!/usr/bin/env python
"""DEBUG: per-item timeline of knn_wunits_kernel for one rank's slice (usage: wu_trace.py WORLD RANK).

Needs a TRACE BUILD of the library: knn_wunits_kernel writing (globaltimer at item start, scan end, merge end; tiles,
units, sub-tile scans) per work item into a __device__ array g_wu_trace[65536 * 4] and an exported
`ksg_debug_wu_trace(unsigned long long* host_out, int n_items)` that copies it out -- the patch is described in
profiles/r2y_wunits_item_trace_slice0of8.txt and is not part of the shipped kernel."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from syntropy_b200 import _engine as E, _cabi, workloads
from syntropy_b200.knn.shannon import _xy_masks, mi1_program
world, rank = int(sys.argv[1]), int(sys.argv[2])
data, call = workloads.c2_mi_1m()
rows = data[call["idxs_x"] + call["idxs_y"], :]
pts = E.to_device_points(rows); n, k = pts.shape[1], call["k"]
masks = _xy_masks(call["idxs_x"], call["idxs_y"]); c = -(-n // world)
q0, q1 = rank * c, min(n, (rank + 1) * c)
prog = mi1_program(E.psi_host(k), E.psi_host(n))
for it in range(4):
    status = E.Status(pts.device)
    E._ksg1_step_ext(pts, k, masks, prog, n, q0, q1, True, None, status); torch.cuda.synchronize()
lib = _cabi.load()._lib if hasattr(_cabi.load(), "_lib") else C.CDLL(os.path.join(ROOT, "syntropy_b200/lib/libksg_b200.so"))
n_items = min(65536, 4 * 3 * (-(-(q1 - q0) // 128)))
buf = (C.c_ulonglong * (4 * n_items))()
if not hasattr(lib, "ksg_debug_wu_trace"):
    raise SystemExit("this library is not a trace build (no ksg_debug_wu_trace): see the docstring")
lib.ksg_debug_wu_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", lib.ksg_debug_wu_trace(buf, n_items))
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4)
live = a[:, 0] > 0
a = a[live]
t0 = a[:, 0].min()
st, en, mg = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3
tiles, nun, subs = (a[:, 3] >> np.uint64(48)).astype(int), ((a[:, 3] >> np.uint64(32)) & np.uint64(0xffff)).astype(int), (a[:, 3] & np.uint64(0xffffffff)).astype(int)
dur = en - st
print(f"items {len(a)}  kernel span {mg.max():.1f} us  item dur: mean {dur.mean():.1f} p50 {np.median(dur):.1f} p90 {np.percentile(dur,90):.1f} p99 {np.percentile(dur,99):.1f} max {dur.max():.1f}")
print(f"start times: p50 {np.median(st):.1f} p90 {np.percentile(st,90):.1f} max {st.max():.1f}")
o = np.argsort(-mg)[:12]
print("last finishers: (start, scan end, merge end, dur, tiles, nunits, subtiles)")
for i in o: print(f"  {st[i]:7.1f} {en[i]:7.1f} {mg[i]:7.1f} {dur[i]:7.1f} {tiles[i]:4d} {nun[i]:4d} {subs[i]:5d}")
o = np.argsort(-dur)[:8]
print("longest items:")
for i in o: print(f"  {st[i]:7.1f} {en[i]:7.1f} {mg[i]:7.1f} {dur[i]:7.1f} {tiles[i]:4d} {nun[i]:4d} {subs[i]:5d}")
print("us per sub-tile scan (items with >= 20 scans): ", np.median(dur[subs >= 20] / subs[subs >= 20]))
print("merge time of merging items: ", np.sort((mg - en)[mg > en])[-8:])
