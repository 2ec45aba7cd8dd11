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
