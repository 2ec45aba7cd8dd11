#!/usr/bin/env python
"""Headline benchmark: KSG-1 mutual information 1D;1D, n = 1e6, k = 5 (BASELINE.json
configs[1], "C2"), samples/sec.

    python bench.py --gpus 1 --steps K --warmup W                 this repo's arm
    torchrun ... bench.py --gpus N ...                            N ranks, query rows sharded
    python bench.py --impl reference ...                          CPU arm (scipy cKDTree port)

One "step" = one full pass of the hot path over the n-sample data set: joint k-NN radius
-> strict marginal counts -> psi pointwise sum -> all-gather -> deterministic mean.
`value` times the step with the point set already resident in HBM; `e2e` times the public
call ``syntropy_b200.knn.mutual_information`` on HOST numpy arrays (H2D of the data and
D2H of the pointwise vector + average inside the timed region).  Rank 0 prints ONE JSON
line.  See DESIGN.md "Measurement" for the roofline accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "KSG MI query samples/sec at n=1M,k=5"
UNIT = "samples/s"
# algorithmic FP-pipe lane-ops per (query, candidate) pair, SURVEY.md 8(d):
# pass 1 (radius) = D FADD + (D-1) FMNMX + 1 FSETP = 2D; pass 2 (counts) = D FADD + M FMNMX + S FSETP
OPS_PER_PAIR = {"c1": (4, 4), "c2": (4, 4), "c3": (12, 14), "c4": (16, 42), "c5": (16, 16)}
CONFIG_SIZES = {"c1": 5_000, "c3": 1_000_000, "c4": 250_000, "c5": 4_000_000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="samples (default: the metric's n = 1e6)")
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--engine", default="windowed", choices=["windowed", "dense", "exact"],
                    help="windowed: fp32 filter + window pruning (default); dense: same kernels over all "
                         "pairs (brute-force roofline case); exact: float64 dense kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense-leg", action="store_true",
                    help="skip the extra brute-force (dense engine) measurement reported under 'dense_engine'")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000,
                    help="samples of the C2 generator the in-line CPU baseline runs on (default: the full n)")
    ap.add_argument("--configs", default="c1,c3,c4,c5",
                    help="other BASELINE configurations measured + parity-checked into the line's 'configs' block "
                         "('' or 'none': skip)")
    ap.add_argument("--config-scale", type=float, default=1.0, help="scale the sizes of the 'configs' block")
    ap.add_argument("--parity-queries", type=int, default=256)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # timing rule: at least three untimed steps (both arms)
    return args


def config_dict(n, k, gpus, engine="windowed"):
    """The `config` object of the JSON line -- the same for both arms (the driver compares them)."""
    return {"workload": f"C2: KSG-1 MI 1D;1D n={n} k={k}, x vs x^2+noise (BASELINE configs[1])",
            "parallelism": f"b200 arm: query rows sharded x{gpus}, full point set per GPU; "
                           "reference arm: host threads on rank 0 (scipy cKDTree workers=-1)",
            "l2": "b200 arm: flushed between timed steps (256 MiB write); reference arm: CPU, not applicable",
            "engine": f"b200 arm: {engine}; reference arm: the reference's algorithm on CPU (oracle/ref_port.py)"}


def workload(n):
    from syntropy_b200 import workloads

    data, call = workloads.c2_mi_1m(n=n, seed=0)
    return data, call


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from
    a thread every ~2 ms (a step of this path is a few ms, too short for `nvidia-smi -lms`);
    if NVML cannot be loaded, one `nvidia-smi` query loop is used instead."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.samples = []      # (sm_mhz, reasons bitmask or None, power_w)
        self.sm_max = None
        self.lines = []
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.source = None

    # ---- NVML
    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        if self.uuid:
            for cand in (f"GPU-{self.uuid}", str(self.uuid)):
                try:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(cand)
                except Exception:
                    continue
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll_nvml(self, pynvml, h):
        while not self.stop_flag.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((float(sm), int(rs), pw))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            pynvml, h = self._nvml_handle()
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._pynvml = pynvml
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, h), daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.source == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            p = self._pynvml
            bits = {"hw_slowdown": p.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": p.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": p.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": p.nvmlClocksEventReasonSwPowerCap}
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"], "source": "nvml"}
            reasons = sorted(name for name, bit in bits.items() if any(rs & bit for _, rs, _ in self.samples))
            powers = [pw for _, _, pw in self.samples if pw is not None]
            return {"sm_mhz": statistics.median(sm for sm, _, _ in self.samples), "sm_max_mhz": self.sm_max,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml",
                    "power_w_max": max(powers) if powers else None}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": "nvidia-smi"}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------ CPU arms
def cpu_port_run(data, call, mode):
    from oracle import ref_port

    t0 = time.perf_counter()
    ptw, avg = ref_port.mutual_information_1(call["idxs_x"], call["idxs_y"], call["k"], data, mode=mode)
    return time.perf_counter() - t0, float(avg)


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm for this path on the host cores.
    The reference itself is pure Python on scipy and cannot travel to the GPU box
    (/root/reference does not exist there); oracle/ref_port.py issues the same scipy
    cKDTree calls, here in its all-threads mode (workers=-1, vectorised ball counts)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_port

    data, call = workload(args.n)
    mode = "fast"
    for _ in range(args.warmup):
        cpu_port_run(data, call, mode)
    times = []
    for _ in range(args.steps):
        dt, avg = cpu_port_run(data, call, mode)
        times.append(dt)
    total = sum(times)
    value = args.n * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args.n, args.k, args.gpus, args.engine),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref_port.cores_used(mode), "kind": "port",
                         "sample": f"full workload per step (n={args.n}); scipy cKDTree, workers=-1, vectorised counts"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "mi_nats": avg,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------ the other BASELINE configurations
def _config_plan(name, data, call):
    """(joint rows, [(estimator label, masks, psi program)], parity subspaces, public call)."""
    from syntropy_b200 import _engine as E
    from syntropy_b200 import knn
    from syntropy_b200.knn import multivariate_mi as mv
    from syntropy_b200.knn.shannon import _xy_masks, cmi_spec, mi1_program

    dim = data.shape[0]
    if name in ("c1", "c5"):
        rows = data[call["idxs_x"] + call["idxs_y"], :]
        masks = _xy_masks(call["idxs_x"], call["idxs_y"])
        ests = [("mutual_information", masks, mi1_program)]
        api = lambda: knn.mutual_information(call["idxs_x"], call["idxs_y"], call["k"], data)  # noqa: E731
    elif name == "c3":
        rows = data[call["idxs_x"] + call["idxs_y"] + call["idxs_z"], :]
        masks, prog = cmi_spec(len(call["idxs_x"]), len(call["idxs_y"]), len(call["idxs_z"]))
        ests = [("conditional_mutual_information", masks, prog)]
        api = lambda: knn.conditional_mutual_information(call["idxs_x"], call["idxs_y"], call["idxs_z"],  # noqa: E731
                                                         call["k"], data)
    else:
        rows = data
        ests = [(fn,) + mv.spec(fn, dim) for fn in ("o_information", "total_correlation",
                                                    "dual_total_correlation", "s_information")]
        api = lambda: knn.o_information(data, call["k"])  # noqa: E731
    parity_masks = ests[0][1]  # (C4: the 16 big / small marginals of the O-information)
    subs = [tuple(d for d in range(rows.shape[0]) if (m >> d) & 1) for m in parity_masks]
    return np.ascontiguousarray(rows), ests, parity_masks, subs, api


def run_config_block(args, world, rank, dev, barrier):
    """C1, C3, C4 (O / TC / DTC / S) and C5 at their BASELINE sizes on the same ranks as the headline:
    device time of every estimator step with the point set resident (CUDA events, max over ranks),
    one end-to-end public call on host arrays, executed pairs and FP-pipe fraction of the radius +
    count passes, and a parity verdict -- radii and EVERY count vector of a random query subset
    bit-exact against the C oracle run on the full point set, the average against the committed
    full-size golden value (tests/golden/full_size.json) at 1e-12 relative."""
    import torch
    import torch.distributed as td

    from syntropy_b200 import _engine as E
    from syntropy_b200 import workloads

    names = [c for c in args.configs.split(",") if c and c != "none"]
    if not names:
        return None
    golden = {}
    try:
        golden = json.load(open(os.path.join(ROOT, "tests", "golden", "full_size.json")))
    except Exception:
        pass
    oc = None
    if rank == 0:
        from oracle import ksg_oracle_c as oc

        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        oracle_threads = oc.set_threads(max(1, (os.cpu_count() or 1) // (2 if local_world > 1 else 1)))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = {}
    for name in names:
        n = max(64, int(CONFIG_SIZES[name] * args.config_scale))
        t_gen = time.perf_counter()
        data, call = workloads.CONFIGS[name](n=n, seed=0)
        t_gen = time.perf_counter() - t_gen
        k = call["k"]
        rows, ests, parity_masks, subs, api = _config_plan(name, data, call)
        pts = E.to_device_points(rows)
        rec = {"n": n, "dim": int(rows.shape[0]), "k": k, "estimators": {}}

        def step(masks, prog):
            status = E.Status(dev)
            pw = E.ksg1_device(pts, k, masks, prog, status=status)
            full, total = E.finish_device(pw, n, status=status)
            pw, full, total = E.settle(pw, n, status, full, total)
            return full, total

        reps = 1 if name == "c5" else 2
        for label, masks, prog in ests:
            step(masks, prog)  # warm-up (graphs, psi table)
            barrier()
            a = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
            b = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
            for i in range(reps):
                a[i].record()
                full, total = step(masks, prog)
                b[i].record()
            barrier()
            t = torch.tensor([min(x.elapsed_time(y) for x, y in zip(a, b))], dtype=torch.float64, device=dev)
            if world > 1:
                td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t.item())
            avg = float(total.item()) / n
            est = {"ms": ms, "samples_per_s": n / (ms * 1e-3), "avg_nats": avg}
            g = golden.get(name, {}).get({"mutual_information": "mi", "conditional_mutual_information": "cmi"}.get(label, label))
            if g and golden[name].get("n") == n:
                est["golden_avg"] = g["avg"]
                est["avg_rel_err"] = abs(avg - g["avg"]) / max(abs(g["avg"]), g["mean_abs_ptw"])
                est["avg_matches_golden_1e-12"] = bool(est["avg_rel_err"] <= 1e-12)
            rec["estimators"][label] = est
        first = rec["estimators"][ests[0][0]]
        rec["ms"], rec["samples_per_s"] = first["ms"], first["samples_per_s"]

        # ---- radii + every count vector once more through the stage-level entry (per-stage timers,
        # executed pairs) and the parity check of a query subset on rank 0
        if E._use_fast(rows.shape[0], k + 1):
            E.timer.enabled = True
            E.timer.reset()
            barrier()
            eps, counts, preps = E.radius_and_counts(pts, k, parity_masks)
            barrier()
            stage_ms = {kn: float(np.sum(v)) for kn, v in E.timer.totals_ms().items()}
            E.timer.enabled = False
            per = [(p_.dim, p_.pairs_evaluated()) for p_ in preps]
            p_knn = float(sum(pe[0] for _, pe in per))
            p_cnt = float(sum(pe[1] for _, pe in per))
            a1, a2 = OPS_PER_PAIR[name]
            cnt_ops = a2 * float(per[0][1][1]) + float(sum(2 * d * pe[1] for d, pe in per[1:]))
            peak = sms * 128 * 1.965e9
            rec["stage_ms"] = stage_ms
            rec["executed_pairs"] = {"knn": p_knn, "count": p_cnt, "dense_per_pass": float(n) * n / world}
            if stage_ms.get("fast_knn"):
                rec["frac"] = a1 * p_knn / (stage_ms["fast_knn"] * 1e-3) / peak
            if stage_ms.get("fast_count"):
                rec["frac_count"] = cnt_ops / (stage_ms["fast_count"] * 1e-3) / peak
        else:
            q0, q1 = 0, n
            eps = E.knn_radius(pts, k + 1, 0, n)
            counts = E.count_subspaces(pts, parity_masks, eps, 0, n)
        if rank == 0:
            t0 = time.perf_counter()
            eps_h, cnt_h = eps.cpu().numpy(), counts.cpu().numpy()
            sel = np.sort(np.random.default_rng(1).choice(n, size=min(args.parity_queries, n), replace=False))
            ref_eps = oc.knn_radius(rows, k + 1, queries=rows[:, sel])
            ok_eps = bool(np.array_equal(eps_h[sel], ref_eps))
            ok_cnt = True
            for row, sub in enumerate(subs):
                want = oc.count_within(rows[list(sub), :], ref_eps, queries=rows[list(sub), :][:, sel])
                ok_cnt = ok_cnt and bool(np.array_equal(cnt_h[row, sel], want))
            golden_ok = [e_["avg_matches_golden_1e-12"] for e_ in rec["estimators"].values()
                         if "avg_matches_golden_1e-12" in e_]
            rec["parity"] = bool(ok_eps and ok_cnt and all(golden_ok))
            rec["parity_detail"] = {"queries": int(sel.size), "radii_bit_exact": ok_eps, "counts_bit_exact": ok_cnt,
                                    "count_vectors": len(subs), "averages_vs_golden": golden_ok or None,
                                    "oracle": f"oracle/ksg_oracle.c brute force fp64 on the full point set, "
                                              f"{oracle_threads} threads", "oracle_s": time.perf_counter() - t0}
            if name == "c5" and n == CONFIG_SIZES["c5"]:
                # pointwise values of the committed 50 000-query subset (fast oracle on the full set)
                try:
                    g5 = np.load(os.path.join(ROOT, "tests", "golden", "full_size_c5.npz"))
                    sel5 = g5["sel"].astype(np.int64)
                    ok5 = bool(np.array_equal(eps_h[sel5], g5["eps"]) and np.array_equal(cnt_h[:, sel5], g5["counts"]))
                    rec["parity_detail"]["golden_subset_50k_bit_exact"] = ok5
                    rec["parity"] = bool(rec["parity"] and ok5)
                except Exception as exc:  # fixture missing: say so, do not fail the bench
                    rec["parity_detail"]["golden_subset_50k_bit_exact"] = f"unavailable: {exc}"
        # ---- public API, host arrays in / host arrays out
        barrier()
        t0 = time.perf_counter()
        _, avg_api = api()
        torch.cuda.synchronize()
        rec["e2e_ms"] = 1e3 * (time.perf_counter() - t0)
        rec["e2e_avg_nats"] = float(avg_api)
        rec["host_generate_s"] = t_gen
        rec["fallback_queries"] = dict(E.stats)
        out[name] = rec
        del pts, eps, counts
        torch.cuda.empty_cache()
        barrier()
    return out


# ------------------------------------------------------------------ B200 arm
def run_b200_arm(args):
    import torch
    import torch.distributed as td

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the KSG path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # rank 0 prints exactly one line on stdout: whatever NCCL prints while the communicator
        # comes up (its version banner) is sent to stderr by pointing fd 1 there meanwhile
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            td.init_process_group("nccl", device_id=torch.device("cuda", local))
            td.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    from syntropy_b200 import _engine as E
    from syntropy_b200 import knn
    from syntropy_b200.knn.shannon import mi1_program, _xy_masks

    data, call = workload(args.n)
    n = data.shape[1]
    k = call["k"]
    masks = _xy_masks(call["idxs_x"], call["idxs_y"])
    rows = data[call["idxs_x"] + call["idxs_y"], :]
    dev = torch.device("cuda", local)
    pts = E.to_device_points(rows)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    E.config.engine = args.engine

    last = {}

    def device_step():
        # what knn.mutual_information does after the upload and before the download: the 16-byte
        # status block is read back (and the step recomputed if the k-NN filter flagged queries)
        status = E.Status(dev)
        pw = E.ksg1_device(pts, k, masks, mi1_program, status=status)
        full, total = E.finish_device(pw, n, status=status)
        pw, full, total = E.settle(pw, n, status, full, total)
        last["pw"] = pw
        return full, total

    # the end-to-end leg reads its inputs from page-locked host memory (the same values)
    data_pinned_t = torch.empty(data.shape, dtype=torch.float64, pin_memory=True)
    data_pinned = data_pinned_t.numpy()
    data_pinned[...] = data

    def e2e_step():
        return knn.mutual_information(call["idxs_x"], call["idxs_y"], k, data_pinned)

    def timed(fn, steps, count_launches=False):
        """K steps, each bracketed by its own CUDA events; L2 flushed between steps outside
        the events; device time = sum over steps; max over ranks."""
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        l0 = E.launch_count()
        out = None
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            out = fn()
            ends[i].record()
        barrier()
        launches = E.launch_count() - l0
        ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item()), launches, out

    # warm-up (also builds the psi table and pays first-launch costs)
    for _ in range(args.warmup):
        device_step()
    warm_res = None
    for _ in range(args.warmup):  # (the second call with the same buffers captures the prepare graphs)
        # the previous result stays alive while the next call runs, as in the timed loop below: the result
        # arrays are page-locked blocks of torch's caching host allocator, and the SECOND block would
        # otherwise be allocated (cudaHostAlloc, ~3.5 ms) inside the timed region
        warm_res = e2e_step()
    del warm_res
    barrier()

    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    E.timer.enabled = True
    E.timer.reset()
    ms_total, launches, (full, total) = timed(device_step, args.steps)
    kernel_ms = E.timer.totals_ms()
    E.timer.enabled = False
    clocks = sampler.stop() if rank == 0 else None
    pairs_done = last["pw"].pairs_evaluated() if last.get("pw") is not None else (0, 0)

    # end to end through the public API, host buffers in / host arrays out
    e2e_each = []

    def e2e_wall(steps):
        barrier()
        t0 = time.perf_counter()
        res = None
        for _ in range(steps):
            t1 = time.perf_counter()
            res = e2e_step()   # (returns host arrays: the call has synchronised)
            e2e_each.append(1e3 * (time.perf_counter() - t1))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item()), res

    e2e_s, (ptw_host, avg_host) = e2e_wall(args.steps)

    # brute-force leg: the same filter kernels over ALL n x n pairs (engine "dense"), the case
    # the FP-pipe roofline of BASELINE.json is defined on; reported beside the headline
    dense_leg = None
    if args.engine == "windowed" and not args.no_dense_leg:
        E.config.engine = "dense"
        device_step()
        E.timer.enabled = True
        E.timer.reset()
        d_steps = 3
        d_ms, _, _ = timed(device_step, d_steps)
        d_kernel_ms = E.timer.totals_ms()
        E.timer.enabled = False
        d_tiles = last["pw"].pairs_evaluated()[0]
        E.config.engine = args.engine
        dense_leg = (d_ms, d_steps, statistics.mean(d_kernel_ms["fast_knn"]), d_tiles)

    # host time to ENQUEUE one resident step (no synchronisation inside): what the binding costs
    def enqueue_only():
        status = E.Status(dev)
        pw = E.ksg1_device(pts, k, masks, mi1_program, status=status)
        return E.finish_device(pw, n, status=status)

    host_us = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enqueue_only()
        host_us.append(1e6 * (time.perf_counter() - t0))
        torch.cuda.synchronize()
    host_enqueue_us = statistics.median(host_us)
    stats_c2 = dict(E.stats)
    configs_block = run_config_block(args, world, rank, dev, barrier) if args.engine == "windowed" else None

    if rank == 0:
        value = n * args.steps / (ms_total * 1e-3)
        e2e_value = n * args.steps / e2e_s
        # roofline of the dominant kernel(s): FP-pipe lane-ops, SURVEY.md 8(d)
        nq = E._dist.slice_of(n, 0, world)
        nq = nq[1] - nq[0]
        a1, a2 = OPS_PER_PAIR["c2"]
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        sm_max_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        sm_load = (clocks or {}).get("sm_mhz")
        peak = sms * 128 * sm_max_mhz * 1e6
        mean_ms = {name: statistics.mean(v) for name, v in kernel_ms.items()}
        if args.engine == "exact":
            knn_name, cnt_name = "knn_radius", "count_subspaces"
            pairs_knn = pairs_cnt = float(n) * float(nq)
        else:
            knn_name, cnt_name = "fast_knn", "fast_count"
            pairs_knn, pairs_cnt = (float(v) for v in pairs_done)  # evaluated (query, candidate) pairs, this rank
        knn_ms = mean_ms.get(knn_name, float("nan"))
        cnt_ms = mean_ms.get(cnt_name, 0.0)
        # the HBM-bound leg: the fused tail launch (psi program + sample order + sum: counts in, perm in,
        # scattered values out) on one rank, the psi program alone on several
        hbm_kernel = "psi_finish" if "psi_finish" in mean_ms else "psi_epilogue"
        psi_ms = mean_ms.get(hbm_kernel, float("nan"))
        ach_knn = a1 * pairs_knn / (knn_ms * 1e-3)
        ach_cnt = a2 * pairs_cnt / (cnt_ms * 1e-3) if cnt_ms else 0.0
        dominant = knn_name if knn_ms >= cnt_ms else cnt_name
        ach = ach_knn if dominant == knn_name else ach_cnt
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback"
        if os.path.exists(peaks_file):
            try:
                hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured"
            except Exception:
                pass
        psi_bytes = nq * (4 * len(masks) + 8 + (4 if hbm_kernel == "psi_finish" else 0))
        # dram bytes (read + write) per launch of the dominant kernel, from the committed
        # `ncu --set full` capture of this same command (profiles/traffic.json says which file)
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(n, k, args.gpus, args.engine),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rows.nbytes),
                    "d2h_bytes_per_step": int(8 * n + 16), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "what": "knn.mutual_information(idxs_x, idxs_y, k, data) on a page-locked host array; "
                            "returns host arrays",
                    "ms_each_call": [round(v, 3) for v in e2e_each]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp_pipe", "kernel": dominant, "achieved": ach / 1e12, "peak": peak / 1e12,
                         "unit": "Tlane-op/s", "frac": ach / peak,
                         "traffic": traffic.get(f"{dominant}_{args.engine}_c2") if world == 1 else None,
                         "peak_basis": f"{sms} SMs x 128 lanes x {sm_max_mhz:.0f} MHz (max clock)",
                         "frac_at_load_clock": (ach / (sms * 128 * sm_load * 1e6)) if sm_load else None,
                         "ops_per_pair": {knn_name: a1, cnt_name: a2},
                         "pairs_per_launch": {knn_name: pairs_knn, cnt_name: pairs_cnt},
                         "dense_pairs": float(n) * float(nq),
                         "kernels_ms": mean_ms,
                         "frac_" + knn_name: ach_knn / peak, "frac_" + cnt_name: ach_cnt / peak,
                         "note": ("achieved = SURVEY 8d lane-ops per pair x pairs the filter actually evaluated / stage "
                                  "time. The windowed engine evaluates ~1e-3 of the n^2 pairs and spends most of its "
                                  "issue slots deciding which pairs to skip (planning, box tests, list upkeep), so its "
                                  "fraction is low by construction; 'dense_engine' is the same step with every pair "
                                  "evaluated (the case the FP-pipe roofline is defined on).")
                         if args.engine == "windowed" else None},
            "roofline_hbm": {"bound": "hbm", "kernel": hbm_kernel, "achieved": psi_bytes / (psi_ms * 1e-3) / 1e9,
                             "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                             "frac": psi_bytes / (psi_ms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": traffic.get(hbm_kernel + "_c2")},
            "fallback_queries": stats_c2,
            "mi_nats": float(avg_host),
            "host_enqueue_us_per_step": host_enqueue_us,
            "binding": os.environ.get("KSG_B200_BINDING", "torch") + " (torch C++ extension over the C ABI)"
                       if os.environ.get("KSG_B200_BINDING", "torch") != "ctypes" else "ctypes over the C ABI",
        }
        if configs_block is not None:
            line["configs"] = configs_block
        if dense_leg is not None:
            d_ms, d_steps, d_knn_ms, d_tiles = dense_leg
            d_pairs = float(d_tiles)
            d_ach = a1 * d_pairs / (d_knn_ms * 1e-3)
            line["dense_engine"] = {
                "what": "same step with engine='dense': the fp32 filter scans every candidate tile (brute force)",
                "value": n * d_steps / (d_ms * 1e-3), "unit": UNIT, "ms_per_step": d_ms / d_steps, "steps": d_steps,
                "roofline": {"bound": "fp_pipe", "kernel": "fast_knn", "achieved": d_ach / 1e12, "peak": peak / 1e12,
                             "unit": "Tlane-op/s", "frac": d_ach / peak, "pairs_per_launch": d_pairs,
                             "kernel_ms": d_knn_ms, "traffic": traffic.get("fast_knn_dense_c2")}}
        if not args.no_cpu_baseline and world == 1:
            from oracle import ref_port

            ns = min(args.cpu_sample, n)
            dt, cpu_avg = cpu_port_run(data[:, :ns], call, "faithful")
            line["cpu_baseline"] = {
                "value": ns / dt, "unit": UNIT, "cores": ref_port.cores_used("faithful"), "kind": "port",
                "sample": (f"the full workload, n={ns}, one pass ({dt:.1f} s)" if ns == n else
                           f"first {ns} samples of the same generator, one pass ({dt:.1f} s)") +
                          "; the reference's own scipy call sequence incl. its per-sample Python loop "
                          f"(single-threaded, as the reference is); host has {os.cpu_count()} cores",
                "mi_nats": cpu_avg}
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
