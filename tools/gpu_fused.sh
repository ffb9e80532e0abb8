#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
timeout 600 python -m pytest tests/test_parity.py -x -q -m gpu -k "fused_drop or scheme_ops or c5_shape or latency_and" 2>&1 | tail -3
for shape in c5 c4; do
  for f in 0 1; do
    python tools/quick_ops.py $LIB --shape $shape --batch 1 --reps 200 --only mult_relin relinearize rotate --opt fused_drop=$f
  done
done
