#!/bin/bash
# FP64-assist microbenchmark, then the full GPU session of gpu_round3.sh
TAG=${1:-r1f}
mkdir -p gpurun_out/$TAG
timeout 120 ./tools/fp64_assist > gpurun_out/$TAG/fp64_assist.json 2>&1; echo "fp64_assist rc=$?"; cat gpurun_out/$TAG/fp64_assist.json
bash tools/gpu_round3.sh $TAG
