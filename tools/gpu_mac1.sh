#!/bin/bash
for v in tools/_variants_macu2.so tools/_variants_macu4.so tools/_variants_macu6.so; do
  python tools/quick_ops.py $v --shape c5 --batch 1 --reps 200 --only ext_prod mult_relin
  python tools/quick_ops.py $v --shape c4 --batch 1 --reps 200 --only ext_prod mult_relin
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:ext_mac -c 6 python tools/quick_ops.py tools/_variants_macu2.so --shape c5 --batch 1 --only mult_relin --reps 2 --warmup 2 2>&1 | grep -i "ext_mac\|duration"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:ext_mac -c 6 python tools/quick_ops.py tools/_variants_macu6.so --shape c5 --batch 1 --only mult_relin --reps 2 --warmup 2 2>&1 | grep -i "ext_mac\|duration"
