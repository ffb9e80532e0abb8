#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
for b in 12 16 24 32 48 64; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only rescale --opt pair_path=0
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only rescale --opt pair_path=2 pair_mode=1
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only rescale --opt pair_path=2 pair_mode=2
done
