#!/bin/bash
# repeat the GPU parity suite and the cluster-heavy cases; synccheck over the cluster / warp-barrier kernels
OUT=gpurun_out/${1:-soak}
mkdir -p $OUT
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done | tee $OUT/repeat.log
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_parity.py -m gpu -x -q \
    -k "(raw_words and (15 or 14 or 13 or 12)) or scheme_ops_match" > $OUT/synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 $OUT/synccheck.log
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_parity.py -m gpu -x -q \
    -k "(raw_words and (15 or 12)) or batched_ops" > $OUT/initcheck.log 2>&1; echo "initcheck rc=$?"; tail -3 $OUT/initcheck.log
