#!/bin/bash
# ncu of the cluster forms (ks_pair.cuh): one ckks::mult and one rescale per call at the C3 shape — full capture and warm launch list
TAG=${1:-r4t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIB=hehub_b200/libhehub_b200.so
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ks_fan_kernel|ks_drop_kernel" -s 6 -c 3 -o $OUT/prof_pair -f \
    python tools/quick_ops.py $LIB --shape c3 --batch 1 --only mult_relin rescale --reps 1 --warmup 3 > $OUT/ncu_pair.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_pair.ncu-rep --page raw --csv > $OUT/raw_pair.csv 2>/dev/null
rm -f $OUT/prof_pair.ncu-rep
python tools/ncu_summary.py $OUT/raw_pair.csv "one ckks::mult (ks_fan_kernel, ks_drop_kernel) and one rescale (ks_drop_kernel) per call, C3 shape ($TAG)" > $OUT/ncu_pair_c3.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 40 --csv --log-file $OUT/launches_single_ct_warm_c3.csv \
    python tools/quick_ops.py $LIB --shape c3 --batch 1 --only mult_relin rotate rescale --reps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
g++ -std=c++17 -O2 -Ihehub_b200/cpp tools/cpp_api_latency.cpp hehub_b200/libhehub_b200.so -Wl,-rpath,$PWD/hehub_b200 -o tools/cpp_api_latency 2> $OUT/cpp_build.log && tools/cpp_api_latency > $OUT/cpp_api_latency.json; cat $OUT/cpp_api_latency.json
cat $OUT/ncu_pair_c3.md | head -60
