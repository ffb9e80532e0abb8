#!/bin/bash
# A/B of programmatic dependent launch: product library with HEHUB_B200_PDL=1/0, and any variant builds
for lib in hehub_b200/libhehub_b200.so tools/_variants_*.so; do
for pdl in 1 0; do
  [ "$pdl" = 0 ] && [ "$lib" != hehub_b200/libhehub_b200.so ] && continue
  echo "== $lib HEHUB_B200_PDL=$pdl"
  HEHUB_B200_PDL=$pdl python tools/quick_ops.py $lib --shape c3 c5
  HEHUB_B200_PDL=$pdl python tools/quick_ops.py $lib --shape c3 c5 --batch 1 --reps 200 --only mult_relin rescale
  HEHUB_B200_PDL=$pdl python tools/quick_ntt.py $lib --logn 12 15 | tail -2
done; done
