#!/bin/bash
# Round-2 GPU session P (2 GPUs): sharded == single-rank test with the round's final kernels, bench at N=2
# (full line with configs).
set -u
O=gpurun_out/r2p
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > $O/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/pytest_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-dense-leg ) > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo done > $O/done.txt
