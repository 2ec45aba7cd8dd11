#!/bin/bash
# Round-2 GPU session AF (1 GPU): grouped leave-one-out count kernel (sub-tile culling on the second-largest box
# gap): identity with the block-level kernel, C4 with and without it, parity, fuzz.
set -u
O=gpurun_out/r2af
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "grouped_count" ) > $O/pytest_identity.log 2>&1
echo "pytest rc=$?" >> $O/pytest_identity.log
python bench.py --steps 3 --warmup 3 --configs c4 --no-dense-leg --no-cpu-baseline > $O/bench_c4_grouped.json 2> $O/bench_c4_grouped.err
KSG_B200_LOO_GROUPS=0 python bench.py --steps 3 --warmup 3 --configs c4 --no-dense-leg --no-cpu-baseline > $O/bench_c4_block.json 2> $O/bench_c4_block.err
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $O/pytest_parity.log 2>&1
echo "pytest rc=$?" >> $O/pytest_parity.log
( time KSG_B200_LOO_GROUPS=force timeout 600 python tools/fuzz_engines.py --cases 200 --seed 71 ) > $O/fuzz71_forced.log 2>&1
echo done > $O/done.txt
