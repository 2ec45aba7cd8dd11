"""world_size-2 gloo test of the query-sharding host logic (syntropy_b200/dist.py):
slices tile [0, n) in rank order, the gathered vector equals the unsharded one on every
rank, and pair counts all-reduce.  The per-slice compute is stood in for by the CPU
oracle here (tests may do that; the product never does)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from oracle import ksg_oracle as orc
        from syntropy_b200 import dist, workloads

        data, call = workloads.c1_readme_mi(n=n, seed=3)
        assert dist.world() == (rank, world_size)
        # the peer-memory exchange needs NCCL ranks with CUDA devices: on gloo the collectives stay
        assert dist.peer_exchange("ptw", 8 * n, torch.device("cpu")) is None
        q0, q1 = dist.my_slice(n)
        # stand-in for K1 -> K2 -> K3 on this rank's query rows
        eps = orc.knn_radius(data, 6, queries=data[:, q0:q1])  # k+1: self is a candidate
        ptw = np.full(q1 - q0, float(orc.psi_int(5)) + float(orc.psi_int(n)))
        for row in (0, 1):
            c = orc.count_within(data[(row,), :], eps, queries=data[(row,), q0:q1])
            ptw -= orc.psi_int(c + 1)
        full = dist.gather_slices(torch.from_numpy(ptw), n).numpy()
        counts = dist.sum_counts(torch.tensor([q1 - q0, 2 * (q1 - q0)], dtype=torch.int64))
        np.save(os.path.join(out_dir, f"ptw_{rank}.npy"), full)
        np.save(os.path.join(out_dir, f"cnt_{rank}.npy"), counts.numpy())
        dist.configure("off")
        assert dist.world() == (0, 1) and dist.my_slice(n) == (0, n)
    finally:
        td.destroy_process_group()


def test_slices_tile_the_range():
    from syntropy_b200 import dist

    for n in (1, 7, 8, 1000, 1_000_003):
        for ws in (1, 2, 3, 8):
            edges = [dist.slice_of(n, r, ws) for r in range(ws)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            assert all(hi >= lo for lo, hi in edges)


@pytest.mark.parametrize("n", [1001, 2000])
def test_two_rank_gather_matches_unsharded(tmp_path, n):
    port = 29500 + (os.getpid() % 2000) + n % 7
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    from oracle import ksg_oracle as orc
    from syntropy_b200 import workloads

    data, _ = workloads.c1_readme_mi(n=n, seed=3)
    ref = orc.mutual_information((0,), (1,), 5, data)[0][0]
    for rank in (0, 1):
        got = np.load(tmp_path / f"ptw_{rank}.npy")
        assert np.array_equal(got, ref)
        assert list(np.load(tmp_path / f"cnt_{rank}.npy")) == [n, 2 * n]


def test_peer_exchange_layout():
    """Pointer arithmetic of dist.PeerExchange (control block, two payload parities, epochs of multi-push
    exchanges) without any device: the layout the kernels of csrc/p2p.cuh are handed."""
    from syntropy_b200 import dist

    ex = dist.PeerExchange.__new__(dist.PeerExchange)  # (no allocation: fields filled by hand)
    ex.payload_bytes, ex.ptrs, ex.rank, ex.world, ex.epoch, ex.round = 4096, [0x10000, 0x90000, 0x110000], 1, 3, 0, 0
    assert ex.flag_ptrs() == ex.ptrs and ex.my_flags == 0x90000
    assert ex.aux_ptrs(0) == [p + 256 for p in ex.ptrs] and ex.aux_ptrs(1) == [p + 320 for p in ex.ptrs]
    assert ex.my_aux(1) == 0x90000 + 320 and ex.my_ticket == 0x90000 + 512
    assert ex.payload_ptrs(0) == [p + 4096 for p in ex.ptrs] and ex.payload_ptrs(1) == [p + 8192 for p in ex.ptrs]
    # flags (16 x 8 B), aux (2 x 16 x 4 B) and the ticket do not overlap and stay inside the control block
    assert 16 * 8 <= dist.PeerExchange.AUX_OFF and dist.PeerExchange.AUX_OFF + 2 * 64 <= dist.PeerExchange.TICKET_OFF
    assert dist.PeerExchange.TICKET_OFF + 4 <= dist.PeerExchange.CTRL
    # one push per exchange: epochs 1, 2, 3 ... with alternating parity
    assert [ex.begin() for _ in range(3)] == [(1, 1), (2, 0), (3, 1)]
    # a two-row upload publishes epochs last - 1 and last; the parity still alternates per EXCHANGE
    assert ex.begin(pushes=2) == (5, 0) and ex.begin(pushes=2) == (7, 1)


def test_peer_exchange_is_off_without_a_process_group(monkeypatch):
    from syntropy_b200 import dist

    assert dist.world() == (0, 1)
    assert dist.peer_exchange("ptw", 1024, torch.device("cpu")) is None

