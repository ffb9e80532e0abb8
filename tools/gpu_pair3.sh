#!/bin/bash
for v in hehub_b200/libhehub_b200.so tools/_variants_notrig.so tools/_variants_c8.so tools/_variants_c8notrig.so; do
  for b in 1 2 4; do
    timeout 300 python tools/quick_ops.py $v --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate --opt pair_path=2
  done
done
