#!/bin/bash
# Round-2 GPU session F (2 GPUs): sharded == single-rank test, bench at N=1 and N=2 (both arms' JSON).
set -u
O=gpurun_out/r2f
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > $O/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/pytest_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_2gpu.json 2> $O/bench_2gpu.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline ) > $O/bench_1gpu.json 2> $O/bench_1gpu.err
KSG_B200_EARLY_ORDER=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 20 --warmup 5 --configs none --no-dense-leg > $O/bench_2gpu_exact_order.json 2> $O/bench_2gpu_exact_order.err
echo done > $O/done.txt
