#!/bin/bash
# Round-2 GPU session U2 (8 GPUs): C2 at N = 8 and 4 with the peer-memory exchange (and N = 8 with NCCL for comparison).
set -u
O=gpurun_out/r2u2
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 $T --nproc-per-node 8 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 --no-dense-leg --configs none ) > $O/bench_8gpu_p2p.json 2> $O/bench_8gpu_p2p.err
( time KSG_B200_P2P=0 timeout 300 $T --nproc-per-node 8 --master-port 29722 bench.py --gpus 8 --steps 20 --warmup 5 --no-dense-leg --configs none ) > $O/bench_8gpu_nccl.json 2> $O/bench_8gpu_nccl.err
( time timeout 300 $T --nproc-per-node 4 --master-port 29723 bench.py --gpus 4 --steps 20 --warmup 5 --no-dense-leg --configs none ) > $O/bench_4gpu_p2p.json 2> $O/bench_4gpu_p2p.err
echo done > $O/done.txt
