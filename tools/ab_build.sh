#!/bin/bash
# Build an A/B variant of the CUDA library with extra -D flags:  tools/ab_build.sh <name> [-DFLAG ...]
# Output: tools/_variants_<name>.so (git-ignored; travels to the GPU box).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc --threads 4 -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
    hehub_b200/csrc/api.cu hehub_b200/csrc/ops.cu hehub_b200/csrc/tables.cu hehub_b200/csrc/host_pipe.cu hehub_b200/csrc/params.cu hehub_b200/csrc/sweep.cu \
    -o tools/_variants_$name.so
echo built tools/_variants_$name.so
