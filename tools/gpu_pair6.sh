#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
timeout 900 python -m pytest tests/test_parity.py -x -q -m gpu 2>&1 | tail -3
for b in 1 2 4 8 16; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate rescale --opt pair_path=0
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate rescale --opt pair_path=1
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only rescale --opt pair_path=2 pair_mode=1
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only rescale --opt pair_path=2 pair_mode=2
done
