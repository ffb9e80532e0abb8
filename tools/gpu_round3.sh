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
python tools/sass_census.py $OUT/src.csv --kernel "ntt_fwd_fast_kernel<(int)12" --butterflies $((4096*24576)) --stalls > $OUT/sass_census_ntt_fwd_fast12.md
# DRAM bytes of that launch, tied to the kernel sources it was taken from (bench.py reports it only while the hash matches)
python - <<PY
import csv, json, sys
sys.path.insert(0, ".")
from bench import kernel_source_sha16
rows = list(csv.reader(open("$OUT/raw.csv")))
hdr, data = rows[0], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
def gb(name):
    v, u = float(data[ix[name]].replace(",", "")), rows[1][ix[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
json.dump({"kernel": data[ix["Kernel Name"]], "source": "ncu --set full --clock-control none, session $TAG (one launch, 4096 rows of N=4096)",
           "source_sha16": kernel_source_sha16(), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": 16 * 4096 * 4096,
           "note": "writes below the algorithmic 134 MB: the tail of the output is still dirty in the 126 MB L2 when the kernel ends"},
          open("$OUT/ntt_fwd_traffic.json", "w"), indent=1)
PY
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
