#!/bin/bash
mkdir -p gpurun_out/variants
for lib in hehub_b200/libhehub_b200.so tools/_variants_v*.so; do
  timeout 300 python tools/quick_ntt.py $lib 12 2>&1 | tail -3
done | tee gpurun_out/variants/plan12.log
