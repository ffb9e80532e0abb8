#!/bin/bash
# Full GPU session: parity, smoke, bench (both arms), e2e chunk sweep, ncu launch list + full capture of the headline kernel.
TAG=${1:-r1d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/bench_ref.json 2>&1; tail -c 400 $OUT/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --sweep-cts 37 > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full (ntt fwd N=4096)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd_fast_kernel -s 3 -c 1 -o $OUT/prof_ntt_fwd -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --extras 0 > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_ntt_fwd.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/prof_ntt_fwd.ncu-rep --page source --csv > $OUT/src.csv 2>/dev/null
rm -f $OUT/prof_ntt_fwd.ncu-rep   # gpurun_out/ is capped at 64 MiB; the csv pages carry what the summaries need
python tools/ncu_summary.py $OUT/raw.csv "ncu --set full: ntt_fwd_fast_kernel<12, RowsIO<0>> (headline kernel), session $TAG" > $OUT/ncu_ntt_fwd_fast12.md
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
for n,row in d["extras"]["ntt_sweep_L1_batch4096"].items():
    print("   N=%6s ntt %.3e (%.3f)  intt %.3e (%.3f)" % (n,row["ntt"]["per_s"],row["ntt"]["frac_hbm"],row["intt"]["per_s"],row["intt"]["frac_hbm"]))
for k,v in d["extras"].items():
    if k.startswith("c"):
        print("  ",k,{kk:(round(vv["per_s"]),round(vv.get("frac_hbm",0),3)) for kk,vv in v.items() if isinstance(vv,dict) and "per_s" in vv}, {kk:vv for kk,vv in v.items() if not isinstance(vv,dict)})
PY
ls -la $OUT
