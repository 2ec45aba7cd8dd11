#!/bin/bash
# Round-2 GPU session Y (1 GPU): per-quarter merge in the warp-level k-NN kernel: every rank's slice of an 8-rank
# C2 step, the full step, parity, fuzz.
set -u
O=gpurun_out/r2y
mkdir -p $O
rm -f $O/slices_of_8_quarter_merge.txt
for r in 0 1 2 3 4 5 6 7; do
  python tools/slice_step.py --world 8 --rank $r >> $O/slices_of_8_quarter_merge.txt 2>&1
done
python tools/slice_step.py --world 1 >> $O/slices_of_8_quarter_merge.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 43 ) > $O/fuzz43.log 2>&1
python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
echo done > $O/done.txt
