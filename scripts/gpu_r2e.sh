#!/bin/bash
OUT=gpurun_out/${1:-r2e}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -15 $OUT/pytest_gpu.log
for dl in 1 0; do
SRL_DEVICE_LOOP=$dl timeout 600 python bench.py --no-cpu-baseline --steps 30 > $OUT/bench_dl$dl.json 2> $OUT/bench_dl$dl.err; echo "bench dl=$dl rc=$?"; tail -3 $OUT/bench_dl$dl.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_dl$dl.json")); print("dl=$dl value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f launches %d step %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"],d["gpu_launches"],d.get("iekf_step")))
except Exception as e: print("no bench line", e)
PY
done
