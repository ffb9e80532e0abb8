#!/bin/bash
# The reference's application benchmark (bench/benchmarks.cpp, README table) through the C++ mirror on the GPU, and the
# unmodified reference itself (oracle/_ref/ref_tool apibench) on the same box's host, one core.
OUT=gpurun_out/${1:-api}
mkdir -p $OUT
g++ -std=c++17 -O2 -Ihehub_b200/cpp tools/cpp_api_bench.cpp hehub_b200/libhehub_b200.so -Wl,-rpath,$PWD/hehub_b200 -o tools/cpp_api_bench 2> $OUT/build.log || { cat $OUT/build.log; exit 1; }
tools/cpp_api_bench | tee $OUT/cpp_api_bench.jsonl
for set in "12 36 20" "13 43 10" "14 48 5" "15 55 3"; do oracle/_ref/ref_tool apibench $set; done | tee $OUT/ref_api_bench.jsonl
