#!/bin/bash
# One GPU session: parity tests, smoke, bench, integer-pipe microbenchmark, ncu launch list + full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/smoke.log
echo "== int_peak"; timeout 120 ./tools/int_peak > $OUT/int_peak.json 2>&1; cat $OUT/int_peak.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/bench_ref.json 2>&1; tail -c 600 $OUT/bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full (ntt fwd)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd_fast -s 3 -c 2 -o $OUT/prof_ntt_fwd -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --extras 0 > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ls -la $OUT
