#!/bin/bash
for rep in 1 2; do
for v in tools/_variants_pf0.so tools/_variants_pf1.so; do
  timeout 300 python tools/quick_ops.py $v --shape c3 --batch 1 --reps 400 --only mult_relin relinearize rotate --opt pair_path=2
done
done
