#!/bin/bash
# Round-2 GPU session W (2 GPUs): peer-memory exchange (psi push + sharded upload over NVLink P2P stores):
# symmetric-memory probe, sharded == single rank, bench at N=2 with and without it.
set -u
O=gpurun_out/r2w
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29731 tools/probe_symm_mem.py > $O/probe.txt 2>&1
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > $O/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/pytest_multi.log
( time timeout 300 $T --master-port 29732 bench.py --gpus 2 --steps 20 --warmup 5 --no-dense-leg --configs none ) > $O/bench_2gpu_p2p.json 2> $O/bench_2gpu_p2p.err
( time KSG_B200_P2P=0 timeout 300 $T --master-port 29733 bench.py --gpus 2 --steps 20 --warmup 5 --no-dense-leg --configs none ) > $O/bench_2gpu_nccl.json 2> $O/bench_2gpu_nccl.err
echo done > $O/done.txt
