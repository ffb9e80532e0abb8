#!/bin/bash
# ncu --set full of one ckks::mult+relin wave (6 kernels) at the C3 and C5 shapes, plus plain timings
TAG=${1:-r1g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIB=hehub_b200/libhehub_b200.so
RX='regex:tensor_kernel|ext_mac|ntt_fwd_fast_kernel|intt_fast_kernel'
for shape in c3 c5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k "$RX" -s 12 -c 6 -o $OUT/prof_$shape -f \
      python tools/quick_ops.py $LIB --shape $shape --only mult_relin --reps 1 --warmup 2 > $OUT/ncu_$shape.log 2>&1; echo "ncu $shape rc=$?"
  ncu -i $OUT/prof_$shape.ncu-rep --page raw --csv > $OUT/raw_$shape.csv 2>/dev/null
  ncu -i $OUT/prof_$shape.ncu-rep --page source --csv > $OUT/src_$shape.csv 2>/dev/null
  rm -f $OUT/prof_$shape.ncu-rep   # gpurun_out/ is capped at 64 MiB
  python tools/ncu_summary.py $OUT/raw_$shape.csv "ckks::mult+relin wave, shape $shape ($TAG)" > $OUT/summary_$shape.md
done
ls -la $OUT
