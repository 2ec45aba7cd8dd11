"""Query-row sharding over the ranks of a ``torch.distributed`` process group.

The path shards naturally: every rank holds the FULL point set, owns a
contiguous slice of the query rows for radius -> counts -> psi, and only the pointwise
slices travel: one all-gather of n/G float64 per rank plus nothing else -- the average
is re-reduced from the gathered vector in a fixed order on every rank, so the result
does not depend on G.  The pair-count kernels of the temporal estimators all-reduce
two int64.

One process per GPU (``torchrun``); backend ``nccl`` on the B200 box (NVLink 5 /
NVSwitch), ``gloo`` in the CPU tests of the host logic.
"""
from __future__ import annotations

import torch
import torch.distributed as td

_mode = "auto"  # "auto": shard when a process group with world_size > 1 exists; "off": never
_group = None


def configure(mode: str = "auto", group=None) -> None:
    """``mode`` in {"auto", "off"}; ``group`` = process group to shard over (default: WORLD)."""
    global _mode, _group
    if mode not in ("auto", "off"):
        raise ValueError("mode must be 'auto' or 'off'")
    _mode, _group = mode, group


def world() -> tuple[int, int]:
    """(rank, world_size) the path shards over; (0, 1) when not sharding."""
    if _mode == "off" or not td.is_available() or not td.is_initialized():
        return 0, 1
    return td.get_rank(_group), td.get_world_size(_group)


def chunk_len(n: int, world_size: int) -> int:
    return -(-n // world_size)


def slice_of(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous query range of ``rank``: equal chunks of ceil(n/G), the tail clipped."""
    c = chunk_len(n, world_size)
    lo = min(n, rank * c)
    return lo, min(n, lo + c)


def my_slice(n: int) -> tuple[int, int]:
    rank, ws = world()
    return slice_of(n, rank, ws)


def gather_slices(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather the per-rank pointwise slices into the full length-n vector (on every rank)."""
    rank, ws = world()
    if ws == 1:
        return local
    c = chunk_len(n_total, ws)
    padded = local
    if local.numel() != c:
        padded = torch.zeros(c, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
    out = torch.empty(ws * c, dtype=local.dtype, device=local.device)
    if td.get_backend(_group) == "nccl":
        td.all_gather_into_tensor(out, padded.contiguous(), group=_group)
    else:
        parts = [out[i * c:(i + 1) * c] for i in range(ws)]
        td.all_gather(parts, padded.contiguous(), group=_group)
    return out[:n_total]


def sum_counts(local: torch.Tensor) -> torch.Tensor:
    """All-reduce (sum) of integer pair counts."""
    _, ws = world()
    if ws == 1:
        return local
    td.all_reduce(local, op=td.ReduceOp.SUM, group=_group)
    return local


def all_gather_into(out: torch.Tensor, local: torch.Tensor) -> None:
    """``out`` (contiguous, world_size * local.numel() elements) = the ranks' ``local`` chunks in rank order."""
    _, ws = world()
    if ws == 1:
        out.copy_(local)
        return
    if td.get_backend(_group) == "nccl":
        td.all_gather_into_tensor(out, local.contiguous(), group=_group)
    else:
        c = local.numel()
        td.all_gather([out[i * c:(i + 1) * c] for i in range(ws)], local.contiguous(), group=_group)


def any_flag(flag: torch.Tensor) -> torch.Tensor:
    """In-place OR (as a max) of an integer flag over the ranks."""
    _, ws = world()
    if ws > 1:
        td.all_reduce(flag, op=td.ReduceOp.MAX, group=_group)
    return flag


# ------------------------------------------------------------------ peer-memory exchange (NVLink P2P)
class PeerExchange:
    """One symmetric buffer per rank, mapped into every rank's address space (torch symmetric memory:
    CUDA VMM allocations whose handles the ranks exchange once), for the all-gather steps of the sharded
    path that are too small for a library collective (csrc/p2p.cuh):

        [ control, 4096 B | payload, parity 0 | payload, parity 1 ]
        control: uint64 flags[16] at 0, int32 aux[2][16] at 256, int32 ticket at 512

    ``begin()`` opens the next exchange (every rank calls it in the same order) and returns
    ``(epoch, parity)``; producers push into ``payload_ptrs(parity)`` of all ranks and release
    ``flag_ptrs()``; the consumer waits on its own flags for ``epoch``."""

    CTRL = 4096
    AUX_OFF = 256
    TICKET_OFF = 512

    def __init__(self, payload_bytes: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        rank, ws = world()
        if ws > 16:
            raise RuntimeError("PeerExchange supports at most 16 ranks")
        self.payload_bytes = (int(payload_bytes) + 255) // 256 * 256
        self.buf = symm_mem.empty(self.CTRL + 2 * self.payload_bytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        grp = _group if _group is not None else td.group.WORLD
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        td.barrier(group=_group)  # every rank's control block is zero before anybody pushes
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.rank, self.world = rank, ws
        self.epoch = 0
        self.round = 0

    def begin(self, pushes: int = 1) -> tuple[int, int]:
        """Open the next exchange of ``pushes`` producer launches per rank (they publish the epochs
        ``last - pushes + 1 .. last`` in stream order).  Returns ``(last epoch, payload parity)``."""
        self.epoch += pushes
        self.round += 1
        return self.epoch, self.round & 1

    def payload_ptrs(self, parity: int) -> list[int]:
        return [p + self.CTRL + parity * self.payload_bytes for p in self.ptrs]

    def flag_ptrs(self) -> list[int]:
        return list(self.ptrs)

    def aux_ptrs(self, parity: int) -> list[int]:
        return [p + self.AUX_OFF + parity * 64 for p in self.ptrs]

    @property
    def my_flags(self) -> int:
        return self.ptrs[self.rank]

    def my_aux(self, parity: int) -> int:
        return self.ptrs[self.rank] + self.AUX_OFF + parity * 64

    @property
    def my_ticket(self) -> int:
        return self.ptrs[self.rank] + self.TICKET_OFF

    def payload(self, parity: int, dtype: torch.dtype) -> torch.Tensor:
        """This rank's payload of ``parity`` as a 1-D tensor view."""
        a = self.CTRL + parity * self.payload_bytes
        return self.buf[a:a + self.payload_bytes].view(dtype)


_exchanges: dict = {}
_p2p_ok: bool | None = None


def peer_exchange(kind: str, payload_bytes: int, device: torch.device):
    """The cached exchange buffer for (kind, size), or None when peer memory is not available (single
    rank, gloo, ``KSG_B200_P2P=0``, or the symmetric allocation failed on some rank -- the decision is
    agreed between the ranks, so either all of them use it or none does)."""
    import os

    global _p2p_ok
    _, ws = world()
    if ws == 1 or _p2p_ok is False or os.environ.get("KSG_B200_P2P", "1") == "0":
        return None
    if td.get_backend(_group) != "nccl":
        return None
    key = (kind, int(payload_bytes), device.index)
    ex = _exchanges.get(key)
    if ex is not None:
        return ex
    ok = torch.ones(1, dtype=torch.int32, device=device)
    try:
        ex = PeerExchange(payload_bytes, device)
    except Exception as exc:  # noqa: BLE001 -- any failure means "no peer memory here"
        import warnings

        warnings.warn(f"syntropy_b200: peer-memory exchange unavailable ({exc}); using NCCL collectives")
        ex = None
        ok.zero_()
    td.all_reduce(ok, op=td.ReduceOp.MIN, group=_group)
    if int(ok.item()) == 0:
        _p2p_ok = False
        return None
    _p2p_ok = True
    if len(_exchanges) >= 8:  # sizes change from call to call in a sweep: keep the cache small
        _exchanges.pop(next(iter(_exchanges)))
    _exchanges[key] = ex
    return ex

