#!/bin/bash
TAG=${1:-t01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.load(open("$OUT/bench.json")); print("value %.1fM e2e %.1fM ms/step %.3f k1 %.4f clocks %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"],d["clocks"]))
PY
