#!/bin/bash
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== diag small"; timeout 300 python scripts/gpu_diag.py small > $OUT/diag_small.log 2>&1; echo "rc=$?"; grep -E "mismatch|parity|HTH rel|iekf|p diff|cap600|fallbacks|timing" $OUT/diag_small.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
for MB in 2 3 4; do
  echo "== bench minb=$MB"; SRL_K1_MINB=$MB timeout 600 python bench.py --no-cpu-baseline --steps 10 > $OUT/bench_minb$MB.json 2> $OUT/bench_minb$MB.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_minb$MB.json"))
print("minb=$MB value %.1fM e2e %.1fM ms/step %.3f k1_avg_ms %.4f frac %.3f clocks %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["frac"],d["clocks"]))
PY
done
echo "== ncu full k1 minb=4"
SRL_K1_MINB=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_assoc -s 6 -c 2 -f -o $OUT/k1_full_mb4 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_mb4.log 2>&1; echo "rc=$?"
echo "== ncu full k1 minb=3"
SRL_K1_MINB=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_assoc -s 6 -c 2 -f -o $OUT/k1_full_mb3 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_mb3.log 2>&1; echo "rc=$?"
ls -la $OUT
