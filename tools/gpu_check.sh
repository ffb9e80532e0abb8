#!/bin/bash
# parity on the GPU, per-op timings, bench without the CPU arm
TAG=${1:-chk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/quick_ops.py hehub_b200/libhehub_b200.so --shape c3 c4 c5 2>&1 | tee $OUT/quick_ops.log
timeout 900 python bench.py ${BENCH_ARGS:---no-cpu} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
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
