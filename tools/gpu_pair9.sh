#!/bin/bash
LIB=hehub_b200/libhehub_b200.so
for b in 16 20 24 32; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate --opt pair_path=0
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate --opt pair_path=2 pair_mode=2
done
# another shape: the example chain (N=4096, L=2)
cat > /tmp/shape.py <<'PY'
PY
