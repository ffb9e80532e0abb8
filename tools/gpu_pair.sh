#!/bin/bash
# two-launch key switch (ks_pair.cuh) against the wave path: parity subset, then per-call times over batch sizes
set -x
LIB=hehub_b200/libhehub_b200.so
timeout 600 python -m pytest tests/test_parity.py -x -q -m gpu -k "scheme_ops or pair_path or rotate or c3_mult or c5_shape" 2>&1 | tail -5
for shape in c3 c4 c5; do
  for b in 1 2 4 8 16; do
    for pp in 0 2; do
      timeout 300 python tools/quick_ops.py $LIB --shape $shape --batch $b --reps 200 --only mult_relin relinearize rotate --opt pair_path=$pp
    done
  done
done
for tpc in 1 2 3 4 6 12; do
  timeout 300 python tools/quick_ops.py $LIB --shape c5 --batch 1 --reps 200 --only mult_relin --opt pair_path=2 pair_tpc=$tpc
done
for tpc in 1 2 4 8; do
  timeout 300 python tools/quick_ops.py $LIB --shape c4 --batch 1 --reps 200 --only mult_relin --opt pair_path=2 pair_tpc=$tpc
done
for tpc in 1 2 4; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch 1 --reps 200 --only mult_relin --opt pair_path=2 pair_tpc=$tpc
done
