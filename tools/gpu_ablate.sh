#!/bin/bash
# Ablation session: where does the N=4096 transform's time go?  (variants from tools/ab_build.sh)
OUT=gpurun_out/${1:-abl}
mkdir -p $OUT
echo "== int_peak"; timeout 120 ./tools/int_peak > $OUT/int_peak.json 2>&1; cat $OUT/int_peak.json
for v in base twconst nostore noload noio allabl; do
  chk="--nocheck"; [ $v = base ] && chk=""
  timeout 300 python tools/quick_ntt.py tools/_variants_$v.so $chk 2>&1 | tail -2
done | tee $OUT/ablate_hbm.log
echo "== L2-resident (2048 polys, one 64 MiB slab)"
for v in base allabl; do
  chk="--nocheck"; [ $v = base ] && chk=""
  timeout 300 python tools/quick_ntt.py tools/_variants_$v.so $chk --polys 2048 --slab-mib 64 2>&1 | tail -2
done | tee $OUT/ablate_l2.log
echo "== ncu k_bfly baseline"
timeout 300 ncu --set full --clock-control none -k regex:k_bfly -c 1 -o $OUT/prof_bfly -f ./tools/int_peak > $OUT/ncu_bfly.log 2>&1; echo rc=$?
ncu -i $OUT/prof_bfly.ncu-rep --page raw --csv > $OUT/bfly_raw.csv 2>/dev/null
