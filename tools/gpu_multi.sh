#!/bin/bash
# multi-GPU bench the way the driver launches it (one rank per GPU), plus the host-link probe
N=${1:-2}; TAG=${2:-r1j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 python tools/pcie_probe.py > $OUT/pcie_probe.json 2>&1; cat $OUT/pcie_probe.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench rc=$?"; tail -3 $OUT/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $OUT/bench_ref_${N}gpu.json 2> $OUT/bench_ref_${N}gpu.err; echo "ref rc=$?"; tail -c 300 $OUT/bench_ref_${N}gpu.json
python - <<PY
import json
txt=open("$OUT/bench_${N}gpu.json").read()
d=json.loads(txt[txt.index("{"):].splitlines()[0])
e=d.pop("extras",{})
print(json.dumps(d)[:1500])
for k,v in e.items():
    if k.startswith("c"): print(k, json.dumps(v)[:600])
PY
