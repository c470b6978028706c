#!/bin/bash
OUT=gpurun_out/${1:-ab1}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for m in 1 0 1 0; do
SRL_MAPPED_RESULT=$m timeout 300 python bench.py --no-cpu-baseline --steps 50 > $OUT/bench_m$m.json 2> $OUT/bench_m$m.err
python - <<PY
import json
d=json.load(open("$OUT/bench_m$m.json")); print("mapped=$m value %.1fM e2e %.1fM ms/step %.3f k1 %.4f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"]))
PY
done
