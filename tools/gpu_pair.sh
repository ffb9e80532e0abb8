#!/bin/bash
# Cluster forms for a few ciphertexts per call (csrc/ks_pair.cuh) against the wave path: parity, then per-call times over
# batch sizes for the key-switch ops and rescale, and plain transforms of a few rows on the mode-2 plans.
#   gpurun -- 'bash tools/gpu_run.sh TAG "bash tools/gpu_pair.sh"'
LIB=hehub_b200/libhehub_b200.so
timeout 900 python -m pytest tests/test_parity.py -x -q -m gpu -k "scheme_ops or pair_path or fused_drop or rotate or c3_mult or c5_shape or latency_and" 2>&1 | tail -3
for b in 1 2 3 4 6 8 12 16 24 32; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate rescale --opt pair_path=0
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch $b --reps 300 --only mult_relin relinearize rotate rescale --opt pair_path=2
done
for tpc in 0 1 2 4; do
  timeout 300 python tools/quick_ops.py $LIB --shape c3 --batch 1 --reps 300 --only mult_relin rescale --opt pair_tpc=$tpc
done
for logn in 12 13; do
  python tools/quick_few_rows.py $LIB --logn $logn --opt latency2_rows=0
  python tools/quick_few_rows.py $LIB --logn $logn --opt latency2_rows=1000
done
for shape in c4 c5; do
  for f in 0 1; do
    python tools/quick_ops.py $LIB --shape $shape --batch 1 --reps 200 --only mult_relin relinearize rotate --opt fused_drop=$f
  done
done
