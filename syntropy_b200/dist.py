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
