#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
timeout 600 python -m pytest tests/test_parity.py -x -q -m gpu -k "scheme_ops or pair_path or rotate or c3_mult or c5_shape" 2>&1 | tail -3
for b in 1 2 3 4 6 8; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate --opt pair_path=0
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate --opt pair_path=2 pair_mode=1
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate --opt pair_path=2 pair_mode=2
done
python tools/phase_probe.py tools/_variants_phase.so --shape c3 --opt pair_path=2 pair_mode=2
