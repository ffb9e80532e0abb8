#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
timeout 600 python -m pytest tests/test_parity.py -x -q -m gpu -k "scheme_ops or pair_path or rotate or c3_mult or c5_shape" 2>&1 | tail -3
python tools/phase_probe.py tools/_variants_phase.so --shape c3
for shape in c3 c4 c5; do
  for b in 1 2 4; do
    for pp in 0 2; do
      timeout 300 python tools/quick_ops.py $LIB --shape $shape --batch $b --reps 200 --only mult_relin relinearize rotate --opt pair_path=$pp
    done
  done
done
python tools/phase_probe.py tools/_variants_phase.so --shape c5 --opt pair_path=2 pair_tpc=1
