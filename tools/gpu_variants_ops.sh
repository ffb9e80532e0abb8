#!/bin/bash
# scheme-level op timings of every A/B build next to the product library
mkdir -p gpurun_out/variants
for lib in hehub_b200/libhehub_b200.so tools/_variants_*.so; do
  timeout 300 python tools/quick_ops.py $lib --shape ${SHAPES:-c3 c5} --only ${ONLY:-ext_prod mult_relin} 2>&1 | tail -3
done | tee gpurun_out/variants/${TAG:-ops}.log
