#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): every estimator family returns, on
every rank, exactly what a single rank computes (sharding off) -- pointwise bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_multi_gpu.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    from syntropy_b200 import dist, knn, workloads

    cases = []
    d, c = workloads.c1_readme_mi(n=20_011, seed=1)
    cases.append(("mi 1D;1D", lambda: knn.mutual_information(c["idxs_x"], c["idxs_y"], c["k"], d)))
    # above the sharded-upload threshold, odd chunk length (peer-memory upload + psi push: csrc/p2p.cuh)
    dl, cl = workloads.c1_readme_mi(n=200_003, seed=11)
    cases.append(("mi 1D;1D n=200003 (sharded upload)", lambda: knn.mutual_information(cl["idxs_x"], cl["idxs_y"], cl["k"], dl)))
    # decimal-quantised data: the speculative step flags queries on every rank -> collective redo
    dq = np.round(dl[:, :70_001] / 0.03) * 0.03
    cases.append(("mi 1D;1D, multiples of 0.03 (collective redo)", lambda: knn.mutual_information(cl["idxs_x"], cl["idxs_y"], cl["k"], dq)))
    d3, c3 = workloads.c3_cmi(n=30_007, seed=2)
    cases.append(("cmi 2D;2D|2D (own-order marginals)", lambda: knn.conditional_mutual_information(
        c3["idxs_x"], c3["idxs_y"], c3["idxs_z"], c3["k"], d3)))
    d4, c4 = workloads.c4_oinfo(n=12_345, seed=3)
    cases.append(("o-information 8 vars", lambda: knn.o_information(d4, c4["k"])))
    cases.append(("total correlation 8 vars", lambda: knn.total_correlation(d4, c4["k"])))
    d5, c5 = workloads.c5_info_rate(n=25_000, seed=4)
    cases.append(("mi 4D;4D (own-order marginals)", lambda: knn.mutual_information(
        c5["idxs_x"], c5["idxs_y"], c5["k"], d5)))
    cases.append(("entropy 3D", lambda: knn.differential_entropy(d4, 3, idxs=(0, 1, 2))))
    rk = np.random.default_rng(8)
    post, prior = rk.standard_normal((3, 20_003)), 1.2 * rk.standard_normal((3, 17_001)) + 0.1
    cases.append(("KL divergence 3D (cross-set search)", lambda: knn.kullback_leibler_divergence(post, prior, 2)))
    cases.append(("mi 1D;1D, KSG-2", lambda: knn.mutual_information(c["idxs_x"], c["idxs_y"], c["k"], d[:, :6_001], algorithm=2)))
    cases.append(("cmi, p = 2 (Minkowski kernels)", lambda: knn.conditional_mutual_information(
        c3["idxs_x"], c3["idxs_y"], c3["idxs_z"], c3["k"], d3[:, :5_003], p=2)))
    from syntropy_b200 import mixed

    rng = np.random.default_rng(6)
    dm = rng.integers(0, 3, size=(1, 30_011))
    cm = rng.standard_normal((2, 30_011)) * (1.0 + 0.3 * dm) + dm
    cases.append(("mixed mi, 3 classes (per-class entropies)", lambda: mixed.mutual_information(dm, cm, "knn", 4)))
    cases.append(("mixed H(D|C)", lambda: mixed.conditional_entropy(dm, cm, "continuous", "knn", 4)))
    series = workloads.c5_series(n=6_000, seed=5)
    ok = True
    for name, fn in cases:
        dist.configure("auto")
        ptw_s, avg_s = fn()
        dist.configure("off")
        ptw_1, avg_1 = fn()
        same = bool(np.array_equal(ptw_s, ptw_1) and avg_s == avg_1)
        ok = ok and same
        if rank == 0:
            print(f"{name}: sharded x{world} == single rank: {same} (avg {avg_s!r})", flush=True)
    dist.configure("auto")
    se_s = knn.sample_entropy((0,), series)
    dist.configure("off")
    se_1 = knn.sample_entropy((0,), series)
    ok = ok and se_s == se_1
    if rank == 0:
        print(f"sample_entropy: sharded == single rank: {se_s == se_1} ({se_s!r})", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    td.all_reduce(flag)
    td.destroy_process_group()
    if int(flag.item()):
        raise SystemExit("multi-GPU results differ from single-rank results")
    if rank == 0:
        print("MULTI_GPU_OK")


if __name__ == "__main__":
    main()
