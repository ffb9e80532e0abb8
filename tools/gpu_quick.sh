#!/bin/bash
# Short GPU session: parity tests + transform rates for the current build.
OUT=gpurun_out/${1:-q}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python tools/quick_ntt.py hehub_b200/libhehub_b200.so --logn 10 11 12 13 14 15 2>&1 | tee $OUT/quick.log
for f in tools/_variants_*.so; do [ -f "$f" ] && timeout 300 python tools/quick_ntt.py $f ${QUICK_ARGS:---logn 12} 2>&1 | tee -a $OUT/quick.log; done
