#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py: one pass per tool, summaries under gpurun_out/sanitize/
set -u
O=gpurun_out/sanitize
mkdir -p $O
export KSG_B200_GRAPHS=0   # direct launches: the sanitizer then names the failing kernel's call site
( time python tools/sanitize_run.py ) > $O/plain.log 2>&1
for tool in memcheck racecheck synccheck initcheck; do
  ( time timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py ) > $O/$tool.log 2>&1
  echo "rc=$?" >> $O/$tool.log
done
grep -H "ERROR SUMMARY\|SANITIZE_RUN\|rc=" $O/*.log > $O/summary.txt
