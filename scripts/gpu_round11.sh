#!/bin/bash
TAG=${1:-r11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
run() { NAME=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 10 > $OUT/bench_$NAME.json 2> $OUT/bench_$NAME.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$NAME.json"))
    print("$NAME value %.1fM e2e %.1fM ms/step %.3f k1_avg_ms %.4f stream %.0f/s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"],d["streaming"]["sweeps_per_s"]))
except Exception as e: print("$NAME bench failed", e); print(open("$OUT/bench_$NAME.err").read()[-600:])
PY
}
run mb5 SRL_FAST_MINB=5
run mb4 SRL_FAST_MINB=4
run mb6 SRL_FAST_MINB=6
