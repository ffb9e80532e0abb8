#!/bin/bash
# Second-style GPU session: parity, microbench, A/B of the persistent pipeline, ncu of the pipelined kernel.
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== int_peak"; timeout 120 ./tools/int_peak > $OUT/int_peak.json 2>&1; cat $OUT/int_peak.json
echo "== bench pipeline=1"; timeout 900 python bench.py --no-cpu > $OUT/bench_pipe1.json 2> $OUT/bench1.err; echo "rc=$?"; tail -2 $OUT/bench1.err
echo "== bench pipeline=0"; timeout 900 python bench.py --no-cpu --pipeline 0 > $OUT/bench_pipe0.json 2> $OUT/bench0.err; echo "rc=$?"; tail -2 $OUT/bench0.err
TAG=$TAG python - <<'PY'
import json,sys,os
for tag in ("pipe1","pipe0"):
    try:
        d=json.load(open("gpurun_out/%s/bench_%s.json" % (os.environ["TAG"], tag)))
    except Exception as e:
        print(tag, "unreadable", e); continue
    print(tag, "value %.3e  frac %.3f  e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
    for n,row in d["extras"]["ntt_sweep_L1_batch4096"].items():
        print("   N=%6s ntt %.3e (%.3f)  intt %.3e (%.3f)" % (n,row["ntt"]["per_s"],row["ntt"]["frac_hbm"],row["intt"]["per_s"],row["intt"]["frac_hbm"]))
    for k,v in d["extras"].items():
        if k.startswith("c"):
            print("  ",k,{kk:(round(vv["per_s"]),round(vv["frac_hbm"],3)) for kk,vv in v.items() if isinstance(vv,dict) and "per_s" in vv})
PY
echo "== ncu full (ntt fwd pipe)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd_pipe -s 3 -c 1 -o $OUT/prof_ntt_fwd_pipe -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --extras 0 > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ls -la $OUT
