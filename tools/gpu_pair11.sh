#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
for tpc in 0 1 2 4; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch 1 --reps 400 --only mult_relin relinearize rotate rescale --opt pair_tpc=$tpc
done
for tpc in 0 1 2 4; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch 2 --reps 400 --only mult_relin relinearize rotate rescale --opt pair_tpc=$tpc
done
python - <<'PY'
import ctypes
rt = ctypes.CDLL("libcudart.so")
PY
python tools/quick_few_rows.py $LIB --logn 13 --rows 8 9 10 12 14 15 16 --opt latency2_rows=1000
