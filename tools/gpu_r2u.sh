#!/bin/bash
# Round-2 GPU session U (8 GPUs): the driver's launch at N=8 with the final code (C2 + C4 + C5), then N=4.
set -u
O=gpurun_out/r2u
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-dense-leg --configs c1,c3,c4,c5 ) > $O/bench_8gpu.json 2> $O/bench_8gpu.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722 \
    bench.py --gpus 4 --steps 20 --warmup 5 --no-dense-leg --configs none ) > $O/bench_4gpu.json 2> $O/bench_4gpu.err
echo done > $O/done.txt
