#!/bin/bash
# compute-sanitizer over the cluster / distributed-shared-memory transforms and the fused ops
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity.py -m gpu -x -q \
      -k "(raw_words and (15 or 14 or 13 or 12)) or scheme_ops or batched_ops or latency or staged_key or rotate_in_place or golden or pair_path" > $OUT/$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 $OUT/$tool.log
done
