#!/bin/bash
TAG=${1:-r04}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== diag small"; timeout 300 python scripts/gpu_diag.py small > $OUT/diag_small.log 2>&1; echo "rc=$?"; grep -E "mismatch|parity|HTH rel|iekf|p diff|cap600|fallbacks|timing|Error|error" $OUT/diag_small.log | head -20
echo "== memcheck small"; timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_diag.py small > $OUT/memcheck.log 2>&1; grep -E "ERROR SUMMARY" $OUT/memcheck.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
for MB in 4 5 6; do
  echo "== bench fast_minb=$MB"; SRL_FAST_MINB=$MB timeout 600 python bench.py --no-cpu-baseline --steps 10 > $OUT/bench_fmb$MB.json 2> $OUT/bench_fmb$MB.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_fmb$MB.json"))
    print("fmb=$MB value %.1fM e2e %.1fM ms/step %.3f k1_avg_ms %.4f frac %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["frac"]))
except Exception as e: print("bench failed", e)
PY
done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full k1_fast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 3 -c 2 -f -o $OUT/k1_fast_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_fast.log 2>&1; echo "rc=$?"
ls -la $OUT
