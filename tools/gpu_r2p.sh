#!/bin/bash
# Round-2 GPU session P (2 GPUs): sharded == single-rank test with the round's final kernels, bench at N=2
# (full line with configs), grouped-kernel ring depth A/B on one GPU.
set -u
O=gpurun_out/r2p
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > $O/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/pytest_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-dense-leg ) > $O/bench_2gpu.json 2> $O/bench_2gpu.err
for st in 2 3 4; do
  ( time KSG_B200_GU_STAGES=$st python bench.py --steps 4 --warmup 2 --configs c3,c5 --no-dense-leg --no-cpu-baseline ) > $O/bench_cfg_st$st.json 2> $O/bench_cfg_st$st.err
done
echo done > $O/done.txt
