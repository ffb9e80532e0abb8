#!/bin/bash
OUT=gpurun_out/${1:-r4x}
mkdir -p $OUT
LIB=hehub_b200/libhehub_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 30 --csv --log-file $OUT/launches_single_ct_warm_c5.csv \
    python tools/quick_ops.py $LIB --shape c5 --batch 1 --only mult_relin --reps 2 --warmup 2 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/quick_ops.py $LIB --shape c5 --batch 1 --reps 200 --only tensor ext_prod rescale mult_relin relinearize rotate
python tools/quick_ops.py $LIB --shape c4 --batch 1 --reps 200 --only tensor ext_prod rescale mult_relin relinearize rotate
