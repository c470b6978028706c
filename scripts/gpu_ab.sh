#!/bin/bash
# A/B of an env switch: $2 = variable name, values 1 and 0; first a parity subset ($3 = -k expression, "all" = whole suite)
OUT=gpurun_out/${1:-ab}; VAR=${2:-SRL_PDL}
mkdir -p $OUT
if [ "$3" = "all" ]; then timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 $OUT/pytest_gpu.log;
elif [ -n "$3" ]; then timeout 900 python -m pytest tests -q -m gpu -x -k "$3" > $OUT/pytest_gpu.log 2>&1; echo "pytest -k rc=$?"; tail -5 $OUT/pytest_gpu.log; fi
for v in 1 0 1 0; do
for P in 100000 12500; do
env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --steps 40 --points $P > $OUT/bench_${v}_$P.json 2> $OUT/bench_${v}_$P.err; tail -2 $OUT/bench_${v}_$P.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${v}_$P.json")); print("$VAR=$v P=$P value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f timleg %.3f step %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["timing_leg_ms_per_step"],d["iekf_step"]["sm_cycles_sums_to_pose"]))
except Exception as e: print("no bench line", e)
PY
done
done
