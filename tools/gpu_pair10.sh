#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
timeout 900 python -m pytest tests/test_parity.py -x -q -m gpu 2>&1 | tail -3
for logn in 12 13; do
python tools/quick_few_rows.py $LIB --logn $logn --opt latency2_rows=0
python tools/quick_few_rows.py $LIB --logn $logn --opt latency2_rows=1000
done
for b in 1 2 4; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only tensor ext_prod rescale mult_relin relinearize rotate
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only tensor ext_prod rescale mult_relin relinearize rotate --opt pair_path=0
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only tensor ext_prod rescale mult_relin relinearize rotate --opt pair_path=0 latency2_rows=0
done
