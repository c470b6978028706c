#!/bin/bash
TAG=${1:-r08}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_gpu.log
run() { NAME=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 10 > $OUT/bench_$NAME.json 2> $OUT/bench_$NAME.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$NAME.json"))
    print("$NAME value %.1fM e2e %.1fM ms/step %.3f k1_avg_ms %.4f frac %.3f stream %.0f/s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["frac"],d["streaming"]["sweeps_per_s"]))
except Exception as e: print("$NAME bench failed", e); print(open("$OUT/bench_$NAME.err").read()[-600:])
PY
}
run mb5 SRL_FAST_MINB=5
run mb6 SRL_FAST_MINB=6
run mb4 SRL_FAST_MINB=4
echo "== ncu full k1_fast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 3 -c 3 -f -o $OUT/k1_fast_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_fast.log 2>&1; echo "rc=$?"
