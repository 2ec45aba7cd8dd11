"""Host orchestration of the KSG path on one B200 (per rank).

PyTorch is used for device memory, streams and (in ``dist.py``) the NCCL process
group only; every kernel is reached through the C ABI of ``include/ksg_b200.h``.
There is no CPU fallback: without a CUDA device ``require_cuda`` raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _cabi
from . import _text
from . import dist as _dist


_cuda_ready = False  # torch's CUDA state is initialised (then the raw, cheap accessors below are valid)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def require_cuda() -> torch.device:
    global _cuda_ready
    if _cuda_ready and _raw_device is not None:
        return torch.device("cuda", _raw_device())
    if not torch.cuda.is_available():
        raise RuntimeError(
            "syntropy_b200 needs a CUDA device (B200 / sm_100a): the KSG path is CUDA-only "
            "and has no CPU fallback"
        )
    dev = torch.device("cuda", torch.cuda.current_device())  # (initialises CUDA)
    _cuda_ready = True
    return dev


def _stream() -> int:
    """Raw handle of torch's current stream on the current device (every C-ABI call takes one;
    ``torch.cuda.current_stream()`` costs ~10 us of Python per call, this ~0.3 us)."""
    if _cuda_ready and _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _masks(masks):
    arr = (C.c_uint32 * len(masks))(*[int(m) for m in masks])
    return arr


def mask_of(dims) -> int:
    m = 0
    for d in dims:
        m |= 1 << int(d)
    return m


# ------------------------------------------------------------------ device helpers
_psi_tables: dict = {}
_psi_host: dict = {}


def psi_table(n_entries: int) -> torch.Tensor:
    """Device table psi(0..n_entries-1), cached per device and grown geometrically."""
    dev = require_cuda()
    key = dev.index
    cur = _psi_tables.get(key)
    if cur is None or cur.numel() < n_entries:
        size = max(int(n_entries), 1024, 0 if cur is None else 2 * cur.numel())
        t = torch.empty(size, dtype=torch.float64, device=dev)
        _cabi.check(_cabi.load().ksg_psi_table(_ptr(t), size, _stream()), "ksg_psi_table")
        _psi_tables[key] = t
        cur = t
    return cur


def psi_host(n: int) -> float:
    """psi(n) as computed by the device table (so host constants and device terms agree)."""
    dev = require_cuda()
    key = (dev.index, int(n))
    v = _psi_host.get(key)
    if v is None:
        v = float(psi_table(int(n) + 1)[int(n)].item())
        _psi_host[key] = v
    return v


# ------------------------------------------------------------------ host <-> device staging
class RowSelection:
    """``data[idxs, :]`` without the intermediate host copy: the selected rows go straight
    from the caller's array into the pinned staging buffer (see ``to_device_points``)."""

    def __init__(self, data: np.ndarray, idxs):
        self.data = data
        self.idxs = tuple(int(i) for i in idxs)
        if data.ndim != 2:
            raise ValueError("data must be (n_variables, n_samples)")
        self.shape = (len(self.idxs), data.shape[1])

    def row(self, r: int) -> np.ndarray:
        return self.data[self.idxs[r]]


def select_rows(data: np.ndarray, idxs) -> RowSelection:
    return RowSelection(np.asarray(data), idxs)


_pinned: dict = {}   # (device index, "in" | "out") -> pinned uint8 tensor, grown geometrically
_stage_done: dict = {}  # device index -> event recorded after the last H2D copy out of the "in" buffer
_copy_pool = None
_COPY_CHUNK = 1 << 18  # elements per host copy task (2 MiB of float64)


def _pool():
    global _copy_pool
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        import os

        # one process per GPU shares the host cores with its siblings (torchrun sets LOCAL_WORLD_SIZE)
        siblings = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        _copy_pool = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // (2 * siblings))))
    return _copy_pool


def _pinned_buffer(dev: torch.device, kind: str, nbytes: int) -> torch.Tensor:
    key = (dev.index, kind)
    cur = _pinned.get(key)
    if cur is None or cur.numel() < nbytes:
        size = max(int(nbytes), 1 << 20, 0 if cur is None else 2 * cur.numel())
        cur = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        _pinned[key] = cur
    return cur


def _host_copy(dst: np.ndarray, src: np.ndarray):
    """dst[...] = src (1-D, numeric; casts like ``astype(float64)``), split over a few threads:
    numpy releases the GIL in its copy loops, and one core cannot saturate host memory."""
    n = dst.shape[0]
    if n <= 2 * _COPY_CHUNK:
        np.copyto(dst, src, casting="unsafe")
        return
    futs = [_pool().submit(np.copyto, dst[a:a + _COPY_CHUNK], src[a:a + _COPY_CHUNK], "unsafe")
            for a in range(0, n, _COPY_CHUNK)]
    for f in futs:
        f.result()


def _rows_are_page_locked(rows: RowSelection, dim: int) -> bool:
    data = rows.data
    if data.dtype != np.float64 or not data.flags.writeable or data.shape[1] < (1 << 16):
        return False
    try:
        for r in range(dim):
            row = rows.row(r)
            if not row.flags.c_contiguous or not torch.from_numpy(row).is_pinned():
                return False
    except Exception:
        return False
    return True


_copy_streams: dict = {}  # device index -> the stream host-to-device row copies are issued on


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _copy_streams.get(dev.index)
    if s is None:
        s = _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s


def wait_rows(row_ready) -> None:
    """Order the current stream after the row copies of ``to_device_points(..., row_events=True)``."""
    for ev in row_ready or ():
        ev.wait()


def to_device_points(rows, row_events: bool = False):
    """(D, n) array of any real dtype (or a ``RowSelection``) -> contiguous float64 SoA tensor
    on the GPU.  cKDTree converts its input to C-contiguous float64 the same way (integer and
    float32 inputs are legal).

    Each row is converted into a pinned staging buffer by a few host threads and sent with
    its own asynchronous copy, so the DMA of row r overlaps the host conversion of row r+1.

    ``row_events``: the copies are issued on a copy stream and ``(points, events)`` is returned,
    events[r] recorded after row r's copy: the consumer (``PreparedSet``: ksg_prepare_rows) starts
    sorting row r while the rows behind it are still in flight.  The caller must order every
    other use of the tensor after the events (``wait_rows``)."""
    dev = require_cuda()
    if not isinstance(rows, RowSelection):
        arr = np.asarray(rows)
        if arr.ndim != 2:
            raise ValueError("data must be (n_variables, n_samples)")
        rows = RowSelection(arr, range(arr.shape[0]))
    dim, n = rows.shape
    if rows.data.dtype.kind not in "fiub":
        raise TypeError(f"data must be real-valued, got dtype {rows.data.dtype}")
    out = torch.empty((dim, n), dtype=torch.float64, device=dev)
    if dim == 0 or n == 0:
        return (out, []) if row_events else out
    events = []
    main = torch.cuda.current_stream()
    if row_events:
        lane = _copy_stream(dev)
        lane.wait_stream(main)  # `out` may be a recycled block with work pending on the main stream
    else:
        lane = main

    def sent():
        if row_events:
            ev = torch.cuda.Event()
            ev.record(lane)
            events.append(ev)

    with torch.cuda.stream(lane):
        if _rows_are_page_locked(rows, dim):
            # the caller's array is already page-locked float64 (torch pinned memory, cudaHostRegister):
            # the DMA reads it in place, no staging copy
            for r in range(dim):
                out[r].copy_(torch.from_numpy(rows.row(r)), non_blocking=True)
                sent()
        else:
            prev = _stage_done.get(dev.index)
            if prev is not None:
                prev.synchronize()  # the staging buffer is reused: its last copies must have left it
            stage = _pinned_buffer(dev, "in", dim * n * 8)[: dim * n * 8].view(torch.float64).view(dim, n)
            stage_np = stage.numpy()
            for r in range(dim):
                _host_copy(stage_np[r], rows.row(r))
                out[r].copy_(stage[r], non_blocking=True)
                sent()
            ev = torch.cuda.Event()
            ev.record(lane)
            _stage_done[dev.index] = ev
    return (out, events) if row_events else out


def to_device_points_sharded(rows):
    """``to_device_points`` for several ranks: every rank sends only ITS slice of every row over PCIe
    (n / G samples instead of n) and the full rows are assembled by one NVLink all-gather per row --
    the ranks share the host's PCIe / memory bandwidth, NVLink is not shared with anything.
    Returns a (D, n) view (row stride G * ceil(n / G)) of the assembled tensor."""
    dev = require_cuda()
    if not isinstance(rows, RowSelection):
        arr = np.asarray(rows)
        rows = RowSelection(arr, range(arr.shape[0]))
    dim, n = rows.shape
    rank, ws = _dist.world()
    c = _dist.chunk_len(n, ws)
    q0, q1 = _dist.slice_of(n, rank, ws)
    local_rows = RowSelection(rows.data[:, q0:q1], rows.idxs) if q1 > q0 else None
    ex = _dist.peer_exchange("pts", dim * ws * c * 8, dev)
    if ex is not None:
        # peer memory: the uploaded columns are stored straight into every rank's (D, G * c) array by a copy
        # kernel (one launch per row, ksg_p2p_push); the consumer's stream then waits for the G flags
        lib = _cabi.load()
        mine = to_device_points(local_rows) if local_rows is not None else None
        last, parity = ex.begin(pushes=dim)
        dst = (C.c_void_p * ws)(*ex.payload_ptrs(parity))
        flg = (C.c_void_p * ws)(*ex.flag_ptrs())
        nbytes = (q1 - q0) * 8 if mine is not None else 0
        for d in range(dim):
            _cabi.check(lib.ksg_p2p_push(_ptr(mine[d]) if mine is not None else None, nbytes, dst,
                                         (d * ws * c + rank * c) * 8, flg, rank, ws, last - dim + 1 + d, ex.my_ticket,
                                         _stream()),
                        "ksg_p2p_push")
        epoch = last
        _cabi.check(lib.ksg_p2p_wait(ex.my_flags, ws, epoch, None, None, _stream()), "ksg_p2p_wait")
        return ex.payload(parity, torch.float64)[: dim * ws * c].view(dim, ws * c)[:, :n]
    local = torch.zeros((dim, c), dtype=torch.float64, device=dev)
    if local_rows is not None:
        local[:, : q1 - q0].copy_(to_device_points(local_rows))
    full = torch.empty((dim, ws * c), dtype=torch.float64, device=dev)
    for d in range(dim):
        _dist.all_gather_into(full[d], local[d])
    return full[:, :n]


SHARDED_UPLOAD_MIN = 1 << 16  # samples; below, the collectives cost more than the copies they save

PINNED_RESULT_MAX_BYTES = 64 << 20     # one result
PINNED_RESULT_TOTAL_BYTES = 512 << 20  # all results alive at a time
_pinned_live = 0


def _release_pinned(nbytes: int) -> None:
    global _pinned_live
    _pinned_live -= nbytes



def to_host(values: torch.Tensor) -> np.ndarray:
    """Device vector -> numpy array owned by the caller.

    Up to 64 MiB (and while less than 512 MiB of such results are alive) the array is backed by a page-locked block from torch's caching host allocator
    (recycled when the array is garbage-collected): the DMA lands directly in the memory that is
    returned -- no second host copy, no first-touch page faults of a fresh allocation.  Larger
    results go through the reusable staging buffer into ordinary memory."""
    global _pinned_live
    n = values.numel()
    nbytes = n * values.element_size()
    if 0 < nbytes <= PINNED_RESULT_MAX_BYTES and _pinned_live + nbytes <= PINNED_RESULT_TOTAL_BYTES:
        out_t = torch.empty(n, dtype=values.dtype, pin_memory=True)
        out_t.copy_(values.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = out_t.numpy()
        # callers that keep many results (sweeps) must not pile up page-locked memory: the bytes
        # handed out are counted until the array dies; above the cap results land in ordinary memory
        _pinned_live += nbytes
        weakref.finalize(out, _release_pinned, nbytes)
        return out
    stage = _pinned_buffer(values.device, "out", nbytes)[:nbytes].view(values.dtype)
    stage.copy_(values.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out = np.empty(n, dtype=stage.numpy().dtype)
    _host_copy(out, stage.numpy())
    return out


@dataclass
class PsiProgram:
    """The reference's pointwise accumulation, step by step (see ksg_psi_epilogue)."""
    init: float
    psi_n: float = 0.0
    scale: float | None = None
    steps: list = field(default_factory=list)  # (subspace, count_offset, sign, centered, divisor)

    def add(self, subspace, sign, count_offset=1, centered=False, divisor=1.0):
        self.steps.append((subspace, count_offset, sign, 1 if centered else 0, float(divisor)))


# ------------------------------------------------------------------ per-kernel timing hook
class KernelTimer:
    """CUDA-event timing of the C-ABI calls, on the stream they are launched on.
    ``bench.py`` switches it on to get the live duration of the dominant kernels for the
    roofline; it is off (zero overhead) otherwise."""

    def __init__(self):
        self._enabled = False
        self.records = []  # (name, start_event, end_event)

    @property
    def enabled(self):
        return self._enabled

    @enabled.setter
    def enabled(self, on):
        self._enabled = bool(on)
        if _text.enabled() and (on or _text._ext is not None):
            _text.load().set_timing(bool(on))  # the fused steps of the torch extension record their own spans

    def reset(self):
        self.records = []
        if _text.enabled() and _text._ext is not None:
            _text._ext.collect_spans()

    def span(self, name):
        return _Span(self, name)

    def totals_ms(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            out.setdefault(name, []).append(a.elapsed_time(b))
        if _text.enabled() and _text._ext is not None:
            for name, values in _text._ext.collect_spans().items():
                out.setdefault(name, []).extend(values)
        return out


class _Span:
    def __init__(self, timer, name):
        self.timer, self.name = timer, name

    def __enter__(self):
        if self.timer.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        if self.timer.enabled:
            self.b.record(torch.cuda.current_stream())
            self.timer.records.append((self.name, self.a, self.b))
        return False


timer = KernelTimer()


def launch_count() -> int:
    return int(_cabi.load().ksg_launch_count())


# ------------------------------------------------------------------ kernel wrappers
MAX_KPRIME = 128  # KSG_MAX_KPRIME of include/ksg_b200.h


def knn_radius(points: torch.Tensor, kprime: int, q_begin: int, q_end: int,
               queries: torch.Tensor | None = None, want_indices: bool = False):
    lib = _cabi.load()
    if kprime > MAX_KPRIME:
        raise NotImplementedError(
            f"k + 1 = {kprime} exceeds the neighbour-list limit of the CUDA kernels ({MAX_KPRIME}); "
            "there is no CPU fallback")
    dim, n = points.shape
    nq = q_end - q_begin
    eps = torch.empty(nq, dtype=torch.float64, device=points.device)
    idx = torch.empty((nq, kprime), dtype=torch.int64, device=points.device) if want_indices else None
    ws_bytes = lib.ksg_knn_radius_workspace(n, max(nq, 1), dim, kprime)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=points.device)
    with timer.span("knn_radius"):
        rc = lib.ksg_knn_radius(_ptr(points), points.stride(0), n, dim,
                                _ptr(queries), 0 if queries is None else queries.stride(0),
                                q_begin, q_end, kprime, _ptr(eps), _ptr(idx), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "ksg_knn_radius")
    return (eps, idx) if want_indices else eps


def count_subspaces(points: torch.Tensor, masks, eps: torch.Tensor, q_begin: int, q_end: int,
                    strict: bool = True, bias: int = -1, per_subspace_eps: bool = False,
                    queries: torch.Tensor | None = None) -> torch.Tensor:
    lib = _cabi.load()
    dim, n = points.shape
    nq = q_end - q_begin
    counts = torch.empty((len(masks), nq), dtype=torch.int32, device=points.device)
    ws = torch.empty(256, dtype=torch.uint8, device=points.device)
    with timer.span("count_subspaces"):
        rc = lib.ksg_count_subspaces(_ptr(points), points.stride(0), n, dim,
                                     _ptr(queries), 0 if queries is None else queries.stride(0),
                                     q_begin, q_end, _masks(masks), len(masks), _ptr(eps),
                                     nq if per_subspace_eps else 0, 1 if strict else 0, bias,
                                     _ptr(counts), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "ksg_count_subspaces")
    return counts


def psi_epilogue(counts: torch.Tensor, prog: PsiProgram, n_total: int) -> torch.Tensor:
    lib = _cabi.load()
    n_sub, nq = counts.shape
    table = psi_table(n_total + 2)
    steps = (_cabi.PsiStep * max(len(prog.steps), 1))()
    for t, (sub, off, sign, centered, div) in enumerate(prog.steps):
        steps[t] = _cabi.PsiStep(sub, off, sign, centered, div)
    ptw = torch.empty(nq, dtype=torch.float64, device=counts.device)
    with timer.span("psi_epilogue"):
        rc = lib.ksg_psi_epilogue(_ptr(counts), nq, n_sub, steps, len(prog.steps), prog.init, prog.psi_n,
                                  0 if prog.scale is None else 1, 0.0 if prog.scale is None else prog.scale,
                                  _ptr(table), table.numel(), _ptr(ptw), _stream())
    _cabi.check(rc, "ksg_psi_epilogue")
    return ptw


def device_sum(values: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    lib = _cabi.load()
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=values.device)
    ws = torch.empty(int(lib.ksg_sum_workspace(values.numel())), dtype=torch.uint8, device=values.device)
    with timer.span("sum_f64"):
        rc = lib.ksg_sum_f64(_ptr(values), values.numel(), _ptr(out), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "ksg_sum_f64")
    return out


# ------------------------------------------------------------------ fast path (prepared set)
class Config:
    """Engine selection.  ``engine``:
      "windowed" (default) fp32 tile filter + exact fp64 resolution, candidate tiles pruned
                           by their bounding boxes;
      "dense"              the same kernels scanning every tile (the brute-force roofline case);
      "exact"              the float64 dense kernels (anchor / fallback; also used for > 8 dims).
    All three give identical radii and counts."""
    engine = "windowed"
    primary = -1  # -1: Hilbert curve over the per-coordinate ranks; -2: kd order; d >= 0: sort by coordinate d
    kd_min_dim = 7  # point sets of at least this many coordinates are prepared in kd order (when primary < 0)
    own_order = True          # windowed engine: low-dimensional marginals in their own order
    own_order_max_dim = 4     # ... when they have at most this many coordinates
    fused_tail = True         # single rank: psi program + sample order + sum in one launch (ksg_psi_finish)
    lazy_axes = True          # speculative 1-D-marginal step: value buckets instead of the coordinate sorts


config = Config()


def order_for(dim: int) -> int:
    """The `primary` argument of ksg_prepare for a point set of `dim` coordinates."""
    if config.primary >= 0:
        return min(config.primary, dim - 1)
    if config.primary == -2 or (config.kd_min_dim is not None and dim >= config.kd_min_dim):
        return -2
    return -1


class PreparedSet:
    """Host handle of a ksg_prepared: the struct, the owning buffer and typed views."""

    def __init__(self, pts: torch.Tensor, primary: int, row_ready=None):
        """``row_ready``: events of ``to_device_points(..., row_events=True)`` -- row d is sorted as
        soon as its own copy has landed; the current stream is ordered after all of them on return."""
        lib = _cabi.load()
        dim, n = pts.shape
        size = int(lib.ksg_prepare_workspace(n, dim))
        self.buf = torch.empty(size, dtype=torch.uint8, device=pts.device)
        self.c = _cabi.Prepared()
        ready = None
        if row_ready:
            assert len(row_ready) == dim
            ready = (C.c_void_p * dim)(*[int(ev.cuda_event) for ev in row_ready])
        with timer.span("prepare"):
            rc = lib.ksg_prepare_rows(_ptr(pts), pts.stride(0), n, dim, primary, ready, _ptr(self.buf), size,
                                      C.byref(self.c), _stream())
        _cabi.check(rc, "ksg_prepare_rows")
        self.n, self.dim, self.ld = n, dim, int(self.c.ld)
        # the library's side streams may still be at work on this set when the constructor returns (see
        # __del__): nothing they could touch goes back to the caching allocator before that join
        self._pts = pts

    def ensure_axes(self):
        """A set prepared with lazy axes (primary = -3: value buckets instead of the per-coordinate sorts)
        gets its exact sorted coordinates now, in stream order; no-op for an ordinary set."""
        if int(self.c.axes_mode) != 0:
            stats["lazy_axes_built"] += 1
            _cabi.check(_cabi.load().ksg_prepare_axes(C.byref(self.c), _stream()), "ksg_prepare_axes")

    def __del__(self):
        # the exact coordinate sorts of an early-order set run on side streams of the library: order
        # the current stream after them BEFORE the buffer goes back to the caching allocator
        try:
            if getattr(self, "c", None) is not None and self.c.axes_ready and _cuda_ready:
                _cabi.load().ksg_prepare_join(C.byref(self.c), _stream())
        except Exception:
            pass

    @classmethod
    def from_raw(cls, buf: torch.Tensor, struct_bytes: torch.Tensor) -> "PreparedSet":
        """Wrap a prepared set built by the torch extension (owning buffer + the ksg_prepared bytes)."""
        self = cls.__new__(cls)
        self.buf = buf
        self.c = _cabi.Prepared.from_buffer_copy(struct_bytes.numpy().tobytes())
        self.n, self.dim, self.ld = int(self.c.n), int(self.c.dim), int(self.c.ld)
        self.struct_bytes = struct_bytes
        return self

    def _view(self, ptr, nbytes, dtype):
        off = int(ptr) - self.buf.data_ptr()
        return self.buf[off:off + nbytes].view(dtype)

    @property
    def sorted_f64(self) -> torch.Tensor:  # (dim, n) view with row stride ld
        return self._view(self.c.sorted_f64, self.dim * self.ld * 8, torch.float64).view(self.dim, self.ld)[:, :self.n]

    @property
    def perm(self) -> torch.Tensor:
        return self._view(self.c.perm, self.n * 4, torch.int32)

    @property
    def flags(self) -> torch.Tensor:
        return self._view(self.c.flags, 4, torch.int32)

    def pairs_evaluated(self):
        """(query, candidate) pairs evaluated so far by the (k-NN filter, count filter) kernels
        on this prepared set (uint64 device counters)."""
        t = self._view(self.c.flags, 64, torch.int64).cpu()
        return int(t[2]), int(t[3])


KNN_MARGIN = 3          # list entries the filter keeps beyond k' (ksg_fast_knn's default; dense engine)
KNN_MARGIN_WINDOWED = 6  # windowed engine: under tight seeds a query meets ~k' + 2 candidates below its
#                          seed; with k' + 3 slots every other list filled up and paid a replace-the-
#                          maximum sweep per further hit (14 % of the units kernel's instructions,
#                          profiles/r2c_*), with k' + 6 the fill phase (append, no sweep) nearly always suffices
KNN_WIDE_MARGIN = 24    # second pass for tie-saturated (quantised) data


def _lists_fit(dim: int, kcap: int) -> bool:
    """Do neighbour lists of kcap entries fit in shared memory beside the candidate ring?
    (Geometry of the dense filter kernel -- 4 / 2 queries per thread -- which also serves the
    windowed engine's deferred queries; the windowed kernels, one query per thread, need less.)"""
    dp = (dim + 1) & ~1
    queries_per_block = 128 * (4 if dp <= 4 else 2)
    stages = 3 if dp <= 4 else 2
    ring = stages * (512 * dp * 4 + 16 * dp * 8)
    return ring + kcap * queries_per_block * 8 <= 200 * 1024 and kcap <= 64


def default_margin(windowed: bool, kprime: int, dim: int) -> int:
    # (from 5 coordinates up the seeds are loose -- ball volume x10 and more -- and every list fills
    #  whatever its length: longer lists would only cost shared memory and longer sweeps there)
    if windowed and dim <= 4 and kprime + KNN_MARGIN_WINDOWED <= 64 and _lists_fit(dim, kprime + KNN_MARGIN_WINDOWED):
        return KNN_MARGIN_WINDOWED
    return KNN_MARGIN


def _fast_knn_call(prep: PreparedSet, kprime: int, q0: int, q1: int, windowed: bool, q_index=None,
                   want_indices: bool = False, margin: int | None = None, nflag: torch.Tensor | None = None):
    """One ksg_fast_knn_margin launch sequence (nothing is read back).  ``nflag``: a zeroed int32[1]
    device counter to use instead of a fresh one."""
    lib = _cabi.load()
    nq = q1 - q0
    if margin is None:
        margin = default_margin(windowed and q_index is None, kprime, prep.dim)
    dev = prep.buf.device
    eps = torch.empty(nq, dtype=torch.float64, device=dev)
    idx = torch.empty((nq, kprime), dtype=torch.int64, device=dev) if want_indices else None
    flags = torch.empty(nq, dtype=torch.int32, device=dev)
    if nflag is None:
        nflag = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(int(lib.ksg_fast_knn_workspace_margin(prep.n, max(nq, 1), prep.dim, kprime, margin)), 1),
                     dtype=torch.uint8, device=dev)
    with timer.span("fast_knn" if q_index is None else "fast_knn_deferred"):
        rc = lib.ksg_fast_knn_margin(C.byref(prep.c), q0, q1, _ptr(q_index), kprime, margin, 1 if windowed else 0,
                                     _ptr(eps), _ptr(idx), _ptr(flags), _ptr(nflag), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "ksg_fast_knn_margin")
    return eps, idx, flags, nflag


def fast_knn(prep: PreparedSet, kprime: int, q0: int, q1: int, windowed: bool, want_indices: bool = False):
    """Exact k'-th-neighbour radii of sorted positions [q0, q1): fp32 filter + fp64 refine.
    What the first pass cannot settle is escalated: outliers the windowed pass set aside go through
    a dense launch of the same filter; when many queries tie beyond the list margin (quantised
    data) the pass is repeated once with a long list; the rest (and tie-saturated neighbourhoods in
    general) goes to the dense fp64 kernel.
    ``want_indices``: also the (nq, k') neighbour positions in the prepared order (KSG-2)."""
    nq = q1 - q0
    eps, nbr, flags, nflag = _fast_knn_call(prep, kprime, q0, q1, windowed, want_indices=want_indices)
    if nq and int(nflag.item()):
        eps, nbr = _fast_knn_escalate(prep, kprime, q0, q1, windowed, want_indices, eps, nbr, flags)
    return (eps, nbr) if want_indices else eps


def _fast_knn_escalate(prep: PreparedSet, kprime: int, q0: int, q1: int, windowed: bool, want_indices: bool,
                       eps: torch.Tensor, nbr, flags: torch.Tensor):
    """The first pass flagged queries (see ``fast_knn``): settle them; returns the final (eps, nbr)."""
    nq = q1 - q0
    n_tied = int((flags == 1).sum().item())
    # worth it when the dense float64 pass over the tied queries (n_tied x n pairs at ~3e12 pairs/s)
    # would cost more than a second filter pass over all of them (~70 ns per query, measured)
    if (windowed and n_tied * prep.n > 4e5 * nq and _lists_fit(prep.dim, kprime + KNN_WIDE_MARGIN)):
        stats["knn_wide_margin_passes"] += 1
        eps, nbr, flags, _ = _fast_knn_call(prep, kprime, q0, q1, windowed, want_indices=want_indices,
                                            margin=KNN_WIDE_MARGIN)
    deferred = (flags == 2).nonzero().flatten()
    if deferred.numel():
        stats["knn_deferred_queries"] += int(deferred.numel())
        rows = (q0 + deferred).to(torch.int32)
        eps2, nbr2, flags2, _ = _fast_knn_call(prep, kprime, 0, int(rows.numel()), False, q_index=rows,
                                               want_indices=want_indices)
        eps[deferred] = eps2
        flags[deferred] = flags2
        if want_indices:
            nbr[deferred] = nbr2
    idx = (flags == 1).nonzero().flatten()
    if idx.numel():
        stats["knn_fallback_queries"] += int(idx.numel())
        pts = prep.sorted_f64
        queries = pts[:, q0 + idx].contiguous()
        if want_indices:
            eps[idx], nbr[idx] = knn_radius(pts, kprime, 0, int(idx.numel()), queries=queries, want_indices=True)
        else:
            eps[idx] = knn_radius(pts, kprime, 0, int(idx.numel()), queries=queries)
    return eps, nbr


def fast_counts(prep: PreparedSet, masks, eps: torch.Tensor, q0: int, q1: int, windowed: bool,
                strict: bool = True, bias: int = -1) -> torch.Tensor:
    """counts[s][q] for sorted positions [q0, q1): 1-D subspaces by binary search in the
    coordinate's sorted array, the others by one fused fp32 pass + exact band resolution."""
    lib = _cabi.load()
    nq = q1 - q0
    dev = prep.buf.device
    prep.ensure_axes()  # (the count kernels read the exact per-coordinate sorts)
    counts = torch.empty((len(masks), nq), dtype=torch.int32, device=dev)
    multi = [s for s, m in enumerate(masks) if bin(m).count("1") > 1]
    single = [s for s, m in enumerate(masks) if bin(m).count("1") == 1]
    if multi:
        mm = [masks[s] for s in multi]
        out = counts if len(multi) == len(masks) else torch.empty((len(mm), nq), dtype=torch.int32, device=dev)
        flags = torch.empty(nq, dtype=torch.int32, device=dev)
        nflag = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(max(int(lib.ksg_fast_count_workspace(prep.n, max(nq, 1), prep.dim, len(mm))), 1),
                         dtype=torch.uint8, device=dev)
        with timer.span("fast_count"):
            rc = lib.ksg_fast_count(C.byref(prep.c), q0, q1, _masks(mm), len(mm), _ptr(eps), 1 if strict else 0,
                                    bias, 1 if windowed else 0, _ptr(out), _ptr(flags), _ptr(nflag), _ptr(ws),
                                    ws.numel(), _stream())
        _cabi.check(rc, "ksg_fast_count")
        if nq and int(nflag.item()):
            idx = flags.nonzero().flatten()
            stats["count_fallback_queries"] += int(idx.numel())
            pts = prep.sorted_f64
            queries = pts[:, q0 + idx].contiguous()
            out[:, idx] = count_subspaces(pts, mm, eps[idx].contiguous(), 0, int(idx.numel()), strict=strict,
                                          bias=bias, queries=queries)
        if out is not counts:
            for row, s in enumerate(multi):
                counts[s] = out[row]
    if single and nq:
        # all 1-D subspaces in one launch (ksg_axis_count_multi writes consecutive rows)
        axes = (C.c_int32 * len(single))(*[int(masks[s]).bit_length() - 1 for s in single])
        out1 = counts if len(single) == len(masks) else torch.empty((len(single), nq), dtype=torch.int32, device=dev)
        with timer.span("axis_count"):
            rc = lib.ksg_axis_count_multi(C.byref(prep.c), axes, len(single), q0, q1, _ptr(eps), 1 if strict else 0, bias,
                                          _ptr(out1), _stream())
        _cabi.check(rc, "ksg_axis_count_multi")
        if out1 is not counts:
            for row, s in enumerate(single):
                counts[s] = out1[row]
    return counts


def scatter_to_original(values_sorted: torch.Tensor, prep: PreparedSet) -> torch.Tensor:
    out = torch.empty_like(values_sorted)
    rc = _cabi.load().ksg_scatter_f64(_ptr(values_sorted), _ptr(prep.perm), prep.n, _ptr(out), _stream())
    _cabi.check(rc, "ksg_scatter_f64")
    return out


stats = {"knn_fallback_queries": 0, "count_fallback_queries": 0, "knn_deferred_queries": 0,
         "knn_wide_margin_passes": 0, "lazy_axes_built": 0}


def _use_fast(dim: int, kprime: int = 1) -> bool:
    """The fp32-filtered engines serve up to 8 joint dimensions and neighbour lists that fit in
    shared memory beside the candidate ring; everything else runs on the exact fp64 kernels."""
    if config.engine not in ("windowed", "dense", "exact"):
        raise ValueError("config.engine must be 'windowed', 'dense' or 'exact'")
    if config.engine == "exact" or dim > 8:
        return False
    dp = (dim + 1) & ~1
    queries_per_block = 128 * (4 if dp <= 4 else 2)
    stages = 3 if dp <= 4 else 2
    ring = stages * (512 * dp * 4 + 16 * dp * 8)
    lists = (kprime + 3) * queries_per_block * 8
    return ring + lists <= 200 * 1024 and kprime + 3 <= 64


def check_finite(points: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """flag (int32[1], or ``out``: only ever set, never cleared) = 1 if any entry is NaN or infinite."""
    flag = torch.zeros(1, dtype=torch.int32, device=points.device) if out is None else out
    lib = _cabi.load()
    if points.is_contiguous():
        _cabi.check(lib.ksg_check_finite(_ptr(points), points.numel(), _ptr(flag), _stream()), "ksg_check_finite")
        return flag
    # (a (D, n) view of an assembled, padded array: rows with unit column stride)
    assert points.dim() == 2 and points.stride(1) == 1
    for d in range(points.shape[0]):
        _cabi.check(lib.ksg_check_finite(_ptr(points[d]), points.shape[1], _ptr(flag), _stream()), "ksg_check_finite")
    return flag


class Status:
    """16 device bytes that an estimator call reads back once, at the very end, together with its
    result: [0:8) the float64 sum of the pointwise values, [8:12) int32 non-finite-input flag,
    [12:16) int32 number of queries the speculative k-NN pass could not settle (see ``ksg1_device``)."""

    def __init__(self, device):
        self.buf = torch.zeros(8, dtype=torch.int32, device=device)
        self.total = self.buf[0:2].view(torch.float64)
        self.finite = self.buf[2:3]
        self.nflag = self.buf[3:4]
        self.ticket = self.buf[4:5]  # ksg_psi_finish / ksg_scatter_sum_f64: 0 on entry, 0 again on exit

    def read_async(self) -> np.ndarray:
        """Enqueue the copy to page-locked host memory; valid after the stream is synchronised."""
        host = torch.empty(4, dtype=torch.int32, pin_memory=True)
        host.copy_(self.buf[:4], non_blocking=True)
        return host.numpy()


class Pointwise:
    """Device-resident pointwise values of one estimator call on this rank.

    ``full`` False: ``values`` is this rank's query slice -- in the order of ``prep`` (sorted
    positions) when ``prep`` is given, else in the caller's sample order (exact engine).
    ``full`` True: ``values`` already holds all n samples in the caller's order on every rank.
    ``preps`` lists every prepared set the call built (for the pair counters).
    ``redo``: set by a speculative run (``ksg1_device``): the values are final only if the
    ``Status.nflag`` they were computed under reads 0; otherwise ``redo()`` returns the settled
    Pointwise."""

    def __init__(self, values, prep=None, full=False, preps=(), redo=None, summed=False, pushed=None):
        self.values, self.prep, self.full, self.preps, self.redo = values, prep, full, list(preps), redo
        self.summed = summed  # full values whose sum already sits in the Status block (fused tail launch)
        # (exchange, epoch, parity): the slices of ALL ranks are being stored into this rank's peer-memory
        # vector by the ranks' psi kernels (ksg_psi_epilogue_push); `values` is unused then
        self.pushed = pushed
        self.flags_summed = False  # Status.nflag already holds the sum over the ranks (ksg_p2p_wait)

    def pairs_evaluated(self):
        a = b = 0
        for p in self.preps:
            x, y = p.pairs_evaluated()
            a, b = a + x, b + y
        return a, b


def finish_device(pw: Pointwise, n_total: int, out: torch.Tensor | None = None, status: "Status | None" = None):
    """All-gather the pointwise slices (NCCL), put them back into the caller's sample order
    if they were computed in sorted order, and reduce the full vector in a fixed order
    (into ``out``, a float64[1] device view, when given; ``status``: into its ``total``, with its
    ticket -- scatter and sum are then one launch)."""
    ticket = None
    if status is not None:
        out, ticket = status.total, status.ticket
    if pw.full:
        if pw.summed and status is not None:
            return pw.values, out
        ptw_full = pw.values
    elif pw.pushed is not None and status is not None:
        # peer-memory exchange: wait for the G flags (folding the ranks' unsettled-query counters into
        # status.nflag), then sample order + sum -- two launches, no library collective, no host sync
        ex, epoch, parity = pw.pushed
        gathered = ex.payload(parity, torch.float64)[:n_total]
        pw.flags_summed = True
        return _text.load().finish_step(gathered, pw.prep.struct_bytes, out, ticket, ex.my_flags, ex.world, epoch,
                                        ex.my_aux(parity), status.nflag), out
    else:
        if pw.pushed is not None:  # (the slices sit in peer memory; only the status-block path consumes them)
            raise RuntimeError("finish_device: a peer-memory exchange is pending; pass the call's Status")
        ptw_full = _dist.gather_slices(pw.values, n_total)
        if pw.prep is not None:
            if out is not None and _text.enabled() and getattr(pw.prep, "struct_bytes", None) is not None:
                # scatter + sum, one call (one launch with a ticket)
                return _text.load().finish_step(ptw_full, pw.prep.struct_bytes, out, ticket, 0, 0, 0, 0, None), out
            ptw_full = scatter_to_original(ptw_full, pw.prep)
    return ptw_full, device_sum(ptw_full, out=out)


def settle(pw: Pointwise, n_total: int, status: Status, ptw_full: torch.Tensor, total: torch.Tensor):
    """Device-resident callers (bench.py): what ``_finish`` does for a speculative run without the
    copy of the result -- read the status block and recompute if the k-NN pass had flagged queries.
    Returns the final ``(pw, ptw_full, total)``."""
    if pw.redo is None:
        return pw, ptw_full, total
    if not pw.flags_summed:
        _dist.sum_counts(status.nflag)  # unsettled queries of ALL ranks: every rank takes the same branch
    host = status.read_async()
    torch.cuda.current_stream().synchronize()
    if int(host[3]) == 0:
        return pw, ptw_full, total
    pw = pw.redo()
    status.nflag.zero_()
    ptw_full, total = finish_device(pw, n_total, status=status)
    return pw, ptw_full, total


def _finish(pw: Pointwise, n_total: int, status: Status):
    """Gather the pointwise slices of all ranks, reduce deterministically, copy out."""
    ptw_full, _ = finish_device(pw, n_total, status=status)
    if pw.redo is not None and not pw.flags_summed:
        _dist.sum_counts(status.nflag)  # unsettled queries of ALL ranks: every rank takes the same branch
    host = status.read_async()
    ptw = to_host(ptw_full)  # synchronises the stream: the status block has landed too
    if int(host[2]) != 0:
        # what scipy's cKDTree constructor raises, reached from knn/utils.py:29
        raise ValueError("data must be finite, check for nan or inf values")
    if pw.redo is not None and int(host[3]) != 0:
        settled = pw.redo()
        status.nflag.zero_()
        return _finish(settled, n_total, status)
    with np.errstate(all="ignore"):
        avg = host[0:2].view(np.float64)[0] / n_total  # ndarray.mean(): sum / N
    return ptw, avg


# ------------------------------------------------------------------ estimator drivers
def ksg1_family(rows, k: int, masks, prog_builder):
    """Joint radius -> strict counts in every subspace -> psi program -> (ptw, avg).

    ``rows`` is the (D, n) joint-space selection in the reference's concatenation
    order; ``masks`` are subspace bitmasks over those D rows; ``prog_builder(psi_k,
    psi_N)`` returns the PsiProgram.
    """
    n_samples = rows.shape[1] if isinstance(rows, RowSelection) else np.asarray(rows).shape[1]
    if _dist.world()[1] > 1 and n_samples >= SHARDED_UPLOAD_MIN:
        pts, row_ready = to_device_points_sharded(rows), None
    else:
        pts, row_ready = to_device_points(rows, row_events=True)
    status = Status(pts.device)
    # (orders the stream after the row copies)
    pw = ksg1_device(pts, k, masks, prog_builder, row_ready=row_ready, status=status)
    check_finite(pts, out=status.finite)
    return _finish(pw, pts.shape[1], status)


def _dims_of(mask: int):
    return [d for d in range(int(mask).bit_length()) if (int(mask) >> d) & 1]


def own_order_subspaces(masks, dim: int):
    """Indices of the subspaces that are counted in a prepared set of their OWN coordinates.

    A marginal of a few coordinates is pruned far better by a space-filling order of those
    coordinates than by the joint order (whose tiles are wide in any low-dimensional
    projection); each such marginal costs one extra ``ksg_prepare`` (a few sorts) and one
    single-subspace pass.  Marginals that keep most of the joint coordinates (leave-one-out
    families) stay in the fused joint-order pass, which shares one distance evaluation among
    all of them; 1-D marginals need no pass at all (binary search)."""
    if config.engine != "windowed" or not config.own_order:
        return []
    out = []
    for s, m in enumerate(masks):
        d = bin(int(m)).count("1")
        if 1 < d <= config.own_order_max_dim and d < dim:
            out.append(s)
    return out


def radius_and_counts(pts: torch.Tensor, k: int, masks, row_ready=None):
    """Fast engines: joint radii and every subspace's strict counts for ALL n samples, in the
    caller's sample order, on every rank (each rank computes its slice of every pass; the
    slices are all-gathered).  Returns ``(eps (n,), counts (S, n) int32, prepared sets)``."""
    lib = _cabi.load()
    dim, n = pts.shape
    windowed = config.engine == "windowed"
    prep = PreparedSet(pts, order_for(dim), row_ready=row_ready)
    preps = [prep]
    q0, q1 = _dist.my_slice(n)
    eps = fast_knn(prep, k + 1, q0, q1, windowed)
    own = own_order_subspaces(masks, dim)
    counts = torch.empty((len(masks), n), dtype=torch.int32, device=pts.device)

    def to_original(slice_sorted, prep_, out_row):
        full = _dist.gather_slices(slice_sorted, n)
        _cabi.check(lib.ksg_scatter_i32(_ptr(full), _ptr(prep_.perm), n, _ptr(out_row), _stream()), "ksg_scatter_i32")

    joint = [s for s in range(len(masks)) if s not in own]
    if joint:
        c = fast_counts(prep, [masks[s] for s in joint], eps, q0, q1, windowed)
        for row, s in enumerate(joint):
            to_original(c[row], prep, counts[s])
    eps_orig = scatter_to_original(_dist.gather_slices(eps, n), prep)
    for s in own:
        dims = _dims_of(masks[s])
        sub = pts[dims, :].contiguous()
        prep_s = PreparedSet(sub, order_for(len(dims)))
        preps.append(prep_s)
        eps_s = torch.empty(n, dtype=torch.float64, device=pts.device)
        _cabi.check(lib.ksg_gather_f64(_ptr(eps_orig), _ptr(prep_s.perm), n, _ptr(eps_s), _stream()), "ksg_gather_f64")
        c = fast_counts(prep_s, [(1 << len(dims)) - 1], eps_s[q0:q1].contiguous(), q0, q1, True)
        to_original(c[0], prep_s, counts[s])
    return eps_orig, counts, preps


def ksg1_device(pts: torch.Tensor, k: int, masks, prog_builder, row_ready=None,
                status: Status | None = None) -> Pointwise:
    """The device-resident part of ``ksg1_family``.  ``row_ready``: see ``to_device_points``; the
    current stream is ordered after those events when this returns.

    ``status`` (single rank, 1-D marginals only -- the MI of two scalar series): the run is
    SPECULATIVE.  The k-NN filter's count of unsettled queries goes to ``status.nflag`` and is not
    read here; the binary-search counts and the psi pass are enqueued right behind the k-NN
    kernels, so the device never waits for the host in mid-call.  The caller reads the status
    block with the result (``_finish`` / ``settle``) and, in the rare flagged case (tie-saturated
    neighbourhoods, set-aside outliers), calls ``Pointwise.redo``: escalation of the flagged
    queries, then the cheap stages again."""
    dim, n = pts.shape
    prog = prog_builder(psi_host(k), psi_host(n))
    if not _use_fast(dim, k + 1):
        wait_rows(row_ready)
        q0, q1 = _dist.my_slice(n)
        eps = knn_radius(pts, k + 1, q0, q1)
        counts = count_subspaces(pts, masks, eps, q0, q1, strict=True, bias=-1)
        return Pointwise(psi_epilogue(counts, prog, n))
    if not own_order_subspaces(masks, dim):
        # everything is counted in the joint order: keep the slices sorted until the end
        windowed = config.engine == "windowed"
        q0, q1 = _dist.my_slice(n)
        # (several ranks: each runs its slice speculatively; the flag counters are summed over the
        #  ranks before anybody looks at them, so all ranks take the same branch -- see settle / _finish)
        speculative = (status is not None and n > 0 and all(bin(int(m)).count("1") == 1 for m in masks))
        if speculative and _text.enabled():
            return _ksg1_step_ext(pts, k, masks, prog, n, q0, q1, windowed, row_ready, status)
        prep = PreparedSet(pts, order_for(dim), row_ready=row_ready)
        if speculative:
            eps, _, flags, _ = _fast_knn_call(prep, k + 1, q0, q1, windowed, nflag=status.nflag)
            counts = fast_counts(prep, masks, eps, q0, q1, windowed)

            def redo():
                eps2, _ = _fast_knn_escalate(prep, k + 1, q0, q1, windowed, False, eps, None, flags)
                counts2 = fast_counts(prep, masks, eps2, q0, q1, windowed)
                return Pointwise(psi_epilogue(counts2, prog, n), prep=prep, preps=[prep])

            return Pointwise(psi_epilogue(counts, prog, n), prep=prep, preps=[prep], redo=redo)
        eps = fast_knn(prep, k + 1, q0, q1, windowed)
        counts = fast_counts(prep, masks, eps, q0, q1, windowed)
        return Pointwise(psi_epilogue(counts, prog, n), prep=prep, preps=[prep])
    _, counts, preps = radius_and_counts(pts, k, masks, row_ready=row_ready)
    return Pointwise(psi_epilogue(counts, prog, n), full=True, preps=preps)


def _ksg1_step_ext(pts, k, masks, prog: PsiProgram, n, q0, q1, windowed, row_ready, status: Status) -> Pointwise:
    """The speculative 1-D-marginal step as ONE call into the torch extension (ksg_torch.cpp:ksg1_step):
    prepared set, k-NN radii, one binary-search count per marginal and the psi program are enqueued
    by C++ without coming back to Python in between."""
    ext = _text.load()
    dim = pts.shape[0]
    axes = [int(m).bit_length() - 1 for m in masks]
    table = psi_table(n + 2)
    steps = prog.steps
    ready = [int(ev.cuda_event) for ev in row_ready] if row_ready else []
    # curve order with lazy axes (-3): the 1-D counts come from value buckets, no coordinate is sorted; the
    # rare set the buckets cannot settle (tied data) is flagged like an unsettled k-NN query -> redo() below
    primary = order_for(dim)
    if primary == -1 and config.lazy_axes:
        primary = -3
    # one rank owns every query: psi program, sample order and the sum are ONE launch (ksg_psi_finish)
    fused_tail = q0 == 0 and q1 == n and n > 0 and config.fused_tail
    # several ranks: the psi kernel pushes its slice into every rank's vector over NVLink (csrc/p2p.cuh)
    rank, ws = _dist.world()
    ex = _dist.peer_exchange("ptw", ws * _dist.chunk_len(n, ws) * 8, pts.device) if ws > 1 and n > 0 else None
    pushed = None
    p2p_args = ([], [], [], 0, 0, 0)
    if ex is not None:
        epoch, parity = ex.begin()
        pushed = (ex, epoch, parity)
        p2p_args = (ex.payload_ptrs(parity), ex.aux_ptrs(parity), ex.flag_ptrs(), rank, epoch, ex.my_ticket)
    ptw, prep_buf, prep_struct, eps, flags, _counts = ext.ksg1_step(
        pts, k, axes, prog.init, prog.psi_n, prog.scale is not None, 0.0 if prog.scale is None else prog.scale,
        [st[0] for st in steps], [st[1] for st in steps], [st[2] for st in steps], [st[3] for st in steps],
        [st[4] for st in steps], table, q0, q1, windowed, primary, default_margin(windowed, k + 1, dim),
        status.nflag, ready, status.total if fused_tail else None, status.ticket if fused_tail else None, *p2p_args)
    prep = PreparedSet.from_raw(prep_buf, prep_struct)

    def redo():
        eps2, _ = _fast_knn_escalate(prep, k + 1, q0, q1, windowed, False, eps, None, flags)
        counts2 = fast_counts(prep, masks, eps2, q0, q1, windowed)
        return Pointwise(psi_epilogue(counts2, prog, n), prep=prep, preps=[prep])

    if fused_tail:
        return Pointwise(ptw, full=True, preps=[prep], redo=redo, summed=True)
    return Pointwise(ptw, prep=prep, preps=[prep], redo=redo, pushed=pushed)


def ksg2_family(rows, k: int, masks, prog_builder):
    """KSG algorithm 2 (reference: knn/shannon.py:199-265, multivariate_mi.py:118-193): per-subspace
    radii from the k joint neighbours, closed-ball counts, psi(count) without the +1."""
    lib = _cabi.load()
    pts = to_device_points(rows)
    dim, n = pts.shape
    status = Status(pts.device)
    check_finite(pts, out=status.finite)
    q0, q1 = _dist.my_slice(n)
    nq = q1 - q0
    prog = prog_builder(psi_host(k), psi_host(n))
    eps_s = torch.empty((len(masks), nq), dtype=torch.float64, device=pts.device)
    if not _use_fast(dim, k + 1):
        _, idx = knn_radius(pts, k + 1, q0, q1, want_indices=True)
        rc = lib.ksg_subspace_radius_from_neighbours(_ptr(pts), pts.stride(0), n, dim, q0, q1, _ptr(idx), k + 1, 1,
                                                     _masks(masks), len(masks), _ptr(eps_s), _stream())
        _cabi.check(rc, "ksg_subspace_radius_from_neighbours")
        counts = count_subspaces(pts, masks, eps_s, q0, q1, strict=False, bias=-1, per_subspace_eps=True)
        return _finish(Pointwise(psi_epilogue(counts, prog, n)), n, status)

    # fast engines: neighbours from the fp32 filter + exact refine (prepared order); every
    # subspace has its own radius, so every multi-dimensional one gets its own single-subspace pass
    windowed = config.engine == "windowed"
    prep = PreparedSet(pts, order_for(dim))
    preps = [prep]
    _, nbr = fast_knn(prep, k + 1, q0, q1, windowed, want_indices=True)
    rc = lib.ksg_subspace_radius_from_neighbours(C.c_void_p(prep.c.sorted_f64), prep.ld, n, dim, q0, q1, _ptr(nbr),
                                                 k + 1, 1, _masks(masks), len(masks), _ptr(eps_s), _stream())
    _cabi.check(rc, "ksg_subspace_radius_from_neighbours")
    counts = torch.empty((len(masks), n), dtype=torch.int32, device=pts.device)

    def to_original(slice_sorted, prep_, out_row):
        full = _dist.gather_slices(slice_sorted, n)
        _cabi.check(lib.ksg_scatter_i32(_ptr(full), _ptr(prep_.perm), n, _ptr(out_row), _stream()), "ksg_scatter_i32")

    all_mask = (1 << dim) - 1
    for s, m in enumerate(masks):
        dims = _dims_of(m)
        e = eps_s[s].contiguous()
        if len(dims) == 1:
            c = fast_counts(prep, [m], e, q0, q1, windowed, strict=False, bias=-1)
            to_original(c[0], prep, counts[s])
            continue
        if int(m) == all_mask:
            prep_s, e_s = prep, e
        else:
            e_orig = scatter_to_original(_dist.gather_slices(e, n), prep)
            prep_s = PreparedSet(pts[dims, :].contiguous(), order_for(len(dims)))
            preps.append(prep_s)
            e_all = torch.empty(n, dtype=torch.float64, device=pts.device)
            _cabi.check(lib.ksg_gather_f64(_ptr(e_orig), _ptr(prep_s.perm), n, _ptr(e_all), _stream()), "ksg_gather_f64")
            e_s = e_all[q0:q1].contiguous()
        c = fast_counts(prep_s, [(1 << len(dims)) - 1], e_s, q0, q1, windowed, strict=False, bias=-1)
        to_original(c[0], prep_s, counts[s])
    return _finish(Pointwise(psi_epilogue(counts, prog, n), full=True, preps=preps), n, status)


def _kl_entropy_device(pts: torch.Tensor, k: int) -> Pointwise:
    """Kozachenko-Leonenko pointwise entropies of a device-resident (D, n) point set
    (reference: knn/shannon.py:81-87): radius pass + ``ksg_entropy_epilogue``."""
    dim, n = pts.shape
    q0, q1 = _dist.my_slice(n)
    prep = None
    if _use_fast(dim, k + 1):
        prep = PreparedSet(pts, order_for(dim))
        eps = fast_knn(prep, k + 1, q0, q1, config.engine == "windowed")
    else:
        eps = knn_radius(pts, k + 1, q0, q1)
    ptw = torch.empty(q1 - q0, dtype=torch.float64, device=pts.device)
    rc = _cabi.load().ksg_entropy_epilogue(_ptr(eps), q1 - q0, dim, psi_host(k), psi_host(n), _ptr(ptw), _stream())
    _cabi.check(rc, "ksg_entropy_epilogue")
    return Pointwise(ptw, prep=prep, preps=[prep] if prep is not None else [])


def kl_entropy(rows: np.ndarray, k: int):
    pts = to_device_points(rows)
    status = Status(pts.device)
    check_finite(pts, out=status.finite)
    return _finish(_kl_entropy_device(pts, k), pts.shape[1], status)


def kl_entropy_classes(rows: np.ndarray, order: np.ndarray, bounds, k: int, pooled: bool):
    """``knn.differential_entropy`` of every class of a partition of the samples of one (D, n)
    array, and optionally of the pooled sample -- the calling pattern of ``syntropy.mixed`` (one
    call per discrete class on ``continuous_vars[:, mask]``, reference: mixed/shannon.py:82,162,
    241,266).  ``order`` (int64[n]) lists the sample indices class after class, ascending inside a
    class; class g is ``order[bounds[g]:bounds[g + 1]]``.

    The array and ``order`` go to the GPU once; every class is gathered on the device, runs the
    same prepared-set / k-NN / epilogue pipeline as ``kl_entropy`` and is reduced in a fixed order;
    the class results are scattered back into sample order on the device; one download at the end.

    Returns ``(by_sample, pooled_ptw, sums)``: ``by_sample[i]`` the pointwise entropy of sample i
    within its own class, ``pooled_ptw`` the pointwise entropies of the pooled sample (or None),
    ``sums[g]`` the float64 sum over class g (``sums[-1]``: over the pooled sample, if asked)."""
    pts = to_device_points(rows)
    dim, n = pts.shape
    status = Status(pts.device)
    check_finite(pts, out=status.finite)
    n_classes = len(bounds) - 1
    order_dev = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int64)).to(pts.device)
    packed = torch.empty(max(n, 1), dtype=torch.float64, device=pts.device)
    sums = torch.zeros(n_classes + 1, dtype=torch.float64, device=pts.device)
    for g in range(n_classes):
        b0, b1 = int(bounds[g]), int(bounds[g + 1])
        if b1 == b0:
            continue
        sub = pts.index_select(1, order_dev[b0:b1])
        full, _ = finish_device(_kl_entropy_device(sub, k), b1 - b0, out=sums[g:g + 1])
        packed[b0:b1].copy_(full)
    by_sample = torch.empty(n, dtype=torch.float64, device=pts.device)
    if n:
        by_sample.index_copy_(0, order_dev, packed[:n])
    pooled_ptw = None
    if pooled:
        pooled_ptw, _ = finish_device(_kl_entropy_device(pts, k), n, out=sums[n_classes:n_classes + 1])
    host_status = status.read_async()
    sums_host = to_host(sums)
    by_sample_host = to_host(by_sample) if n else np.empty(0)
    pooled_host = to_host(pooled_ptw) if pooled_ptw is not None else None
    if int(host_status[2]) != 0:
        raise ValueError("data must be finite, check for nan or inf values")
    return by_sample_host, pooled_host, sums_host


class CrossSet:
    """Host handle of a ksg_cross: query points that are not part of the prepared candidate set,
    placed on the candidates' space-filling curve (``ksg_cross_prepare``)."""

    def __init__(self, cand: PreparedSet, queries: torch.Tensor):
        lib = _cabi.load()
        dim, nq = queries.shape
        assert dim == cand.dim
        size = int(lib.ksg_cross_prepare_workspace(cand.n, nq, dim))
        self.buf = torch.empty(size, dtype=torch.uint8, device=queries.device)
        self.c = _cabi.Cross()
        with timer.span("cross_prepare"):
            rc = lib.ksg_cross_prepare(C.byref(cand.c), _ptr(queries), queries.stride(0), nq, _ptr(self.buf), size,
                                       C.byref(self.c), _stream())
        _cabi.check(rc, "ksg_cross_prepare")
        self.cand, self.nq, self.dim, self.ld = cand, nq, dim, int(self.c.ld)

    def _view(self, ptr, nbytes, dtype):
        off = int(ptr) - self.buf.data_ptr()
        return self.buf[off:off + nbytes].view(dtype)

    @property
    def perm(self) -> torch.Tensor:
        return self._view(self.c.perm, self.nq * 4, torch.int32)

    @property
    def sorted_f64(self) -> torch.Tensor:
        return self._view(self.c.sorted_f64, self.dim * self.ld * 8, torch.float64).view(self.dim, self.ld)[:, :self.nq]


def fast_knn_cross(cross: CrossSet, kprime: int, q0: int, q1: int) -> torch.Tensor:
    """Exact kprime-th-neighbour distance from the queries at cross positions [q0, q1) to the
    candidate set: windowed fp32 filter + fp64 refine; whatever it sets aside or cannot decide is
    recomputed by the dense float64 kernel."""
    lib = _cabi.load()
    cand = cross.cand
    nq = q1 - q0
    dev = cross.buf.device
    eps = torch.empty(nq, dtype=torch.float64, device=dev)
    flags = torch.empty(nq, dtype=torch.int32, device=dev)
    nflag = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(int(lib.ksg_fast_knn_workspace(cand.n, max(nq, 1), cand.dim, kprime)), 1),
                     dtype=torch.uint8, device=dev)
    with timer.span("fast_knn_cross"):
        rc = lib.ksg_fast_knn_cross(C.byref(cand.c), C.byref(cross.c), q0, q1, kprime, _ptr(eps), _ptr(flags),
                                    _ptr(nflag), _ptr(ws), ws.numel(), _stream())
    _cabi.check(rc, "ksg_fast_knn_cross")
    if nq and int(nflag.item()):
        idx = flags.nonzero().flatten()
        stats["knn_fallback_queries"] += int(idx.numel())
        queries = cross.sorted_f64[:, q0 + idx].contiguous()
        eps[idx] = knn_radius(cand.sorted_f64, kprime, 0, int(idx.numel()), queries=queries)
    return eps


def kl_divergence(post_rows: np.ndarray, prior_rows: np.ndarray, k: int):
    """kNN KL divergence (reference: knn/shannon.py:350-410): rho = (k+1)-th neighbour distance
    inside the posterior sample, nu = k-th neighbour distance of every posterior sample in the
    prior sample (cross-set search on the prior's prepared set)."""
    lib = _cabi.load()
    P = to_device_points(post_rows)
    Q = to_device_points(prior_rows)
    dim, n = P.shape
    m = Q.shape[1]
    status = Status(P.device)
    check_finite(P, out=status.finite)
    check_finite(Q, out=status.finite)
    q0, q1 = _dist.my_slice(n)
    with np.errstate(all="ignore"):
        log_ratio = float(np.log(m / (n - 1)))
    if not _use_fast(dim, k + 1):
        rho = knn_radius(P, k + 1, q0, q1)
        nu = knn_radius(Q, k, q0, q1, queries=P)
        ptw = torch.empty(q1 - q0, dtype=torch.float64, device=P.device)
        rc = lib.ksg_kl_epilogue(_ptr(nu), _ptr(rho), q1 - q0, dim, log_ratio, _ptr(ptw), _stream())
        _cabi.check(rc, "ksg_kl_epilogue")
        return _finish(Pointwise(ptw), n, status)
    windowed = config.engine == "windowed"
    prep_p = PreparedSet(P, order_for(dim))
    rho = scatter_to_original(_dist.gather_slices(fast_knn(prep_p, k + 1, q0, q1, windowed), n), prep_p)
    prep_q = PreparedSet(Q, -1)
    if windowed:
        cross = CrossSet(prep_q, P)
        nu_cross = _dist.gather_slices(fast_knn_cross(cross, k, q0, q1), n)
        nu = torch.empty_like(nu_cross)
        _cabi.check(lib.ksg_scatter_f64(_ptr(nu_cross), _ptr(cross.perm), n, _ptr(nu), _stream()), "ksg_scatter_f64")
    else:
        nu = _dist.gather_slices(knn_radius(Q, k, q0, q1, queries=P), n)
    ptw = torch.empty(n, dtype=torch.float64, device=P.device)
    rc = lib.ksg_kl_epilogue(_ptr(nu), _ptr(rho), n, dim, log_ratio, _ptr(ptw), _stream())
    _cabi.check(rc, "ksg_kl_epilogue")
    return _finish(Pointwise(ptw, full=True, preps=[prep_p, prep_q]), n, status)


def _self_pair_count(templates: torch.Tensor, r: float) -> torch.Tensor:
    """Ordered pairs (i, j), self pairs included, of the columns of ``templates`` (d, N) within
    max-norm distance <= r: this rank's share, as a 1-element int64 tensor.  The template matrix
    is prepared in its own space-filling order and counted by the single-subspace windowed pass
    (closed ball, exact)."""
    lib = _cabi.load()
    d, nt = templates.shape
    prep = PreparedSet(templates, -1)
    q0, q1 = _dist.my_slice(nt)
    out = torch.zeros(1, dtype=torch.int64, device=templates.device)
    if q1 > q0:
        eps = torch.full((q1 - q0,), float(r), dtype=torch.float64, device=templates.device)
        counts = fast_counts(prep, [(1 << d) - 1] if d > 1 else [1], eps, q0, q1, config.engine == "windowed",
                             strict=False, bias=0)
        _cabi.check(lib.ksg_sum_i32(_ptr(counts), counts.numel(), _ptr(out), _stream()), "ksg_sum_i32")
    return out


def _templates(series: torch.Tensor, width: int, nt: int) -> torch.Tensor:
    """(width, nt) matrix of the length-``width`` sliding windows of a device series."""
    return torch.stack([series[w:w + nt] for w in range(width)]).contiguous()


def pair_counts(x: np.ndarray, y: np.ndarray, m: int, r: float):
    """(small, big) closed-ball ordered pair counts of the m / (m+1) templates
    (reference: cKDTree.count_neighbors, knn/temporal.py:71-72,142-143).

    Fast engines, m + 1 <= 8: every count is a self pair count of a prepared template set.  For
    two different series the cross count follows from three self counts,
    ``XY = (UU - XX - YY) / 2`` with U the union of both template sets (the predicate is
    symmetric and all counts are integers, so this is exact).  Otherwise: the dense float64
    kernel."""
    dev = require_cuda()
    xs = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    ys = xs if y is x else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
    nt = xs.numel() - m
    if nt <= 0:
        raise ValueError("series shorter than the template length")
    flag = check_finite(xs) | check_finite(ys)
    if config.engine != "exact" and m + 1 <= 8 and np.isfinite(r) and r >= 0:
        if int(flag.item()):
            raise ValueError("data must be finite, check for nan or inf values")
        res = []
        for width in (m, m + 1):
            tx = _templates(xs, width, nt)
            if ys is xs:
                res.append(_self_pair_count(tx, r))
            else:
                ty = _templates(ys, width, nt)
                uu = _self_pair_count(torch.cat([tx, ty], dim=1).contiguous(), r)
                res.append(uu - _self_pair_count(tx, r) - _self_pair_count(ty, r))
        out = _dist.sum_counts(torch.cat(res))
        if ys is not xs:
            out = out // 2
        host = torch.cat([out, flag.to(torch.int64)]).cpu().numpy()
        if host[2] != 0:
            raise ValueError("data must be finite, check for nan or inf values")
        return int(host[0]), int(host[1])
    return _pair_counts_dense(xs, ys, nt, m, r, flag)


def _pair_counts_dense(xs, ys, nt, m, r, flag):
    dev = xs.device
    i0, i1 = _dist.my_slice(nt)
    out = torch.zeros(2, dtype=torch.int64, device=dev)
    rc = _cabi.load().ksg_paircount_fixed_radius(_ptr(xs), _ptr(ys), nt, m, float(r), i0, i1, _ptr(out), _stream())
    _cabi.check(rc, "ksg_paircount_fixed_radius")
    out = _dist.sum_counts(out)
    host = torch.cat([out, flag.to(torch.int64)]).cpu().numpy()
    if host[2] != 0:
        raise ValueError("data must be finite, check for nan or inf values")
    return int(host[0]), int(host[1])
