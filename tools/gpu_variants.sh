#!/bin/bash
# time every A/B build next to the product library: plain transforms and the scheme-level ops
mkdir -p gpurun_out/variants
for lib in hehub_b200/libhehub_b200.so tools/_variants_*.so; do
  timeout 300 python tools/quick_ntt.py $lib --logn ${LOGNS:-14} 2>&1 | tail -3
  timeout 300 python tools/quick_ops.py $lib --shape ${SHAPES:-c4} 2>&1 | tail -3
done | tee gpurun_out/variants/${TAG:-ab}.log
