#!/bin/bash
# A/B of an env switch: $2 = variable name, values 1 and 0; plus the whole GPU suite first if $3 = tests
OUT=gpurun_out/${1:-ab}; VAR=${2:-SRL_PDL}
mkdir -p $OUT
if [ "$3" = "tests" ]; then timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 $OUT/pytest_gpu.log; fi
for v in 1 0 1 0; do
env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --steps 40 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "bench $VAR=$v rc=$?"; tail -2 $OUT/bench_$v.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$v.json")); print("$VAR=$v value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f timleg %.3f launches %d step %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["timing_leg_ms_per_step"],d["gpu_launches"],d["iekf_step"]["sm_cycles_sums_to_pose"]))
except Exception as e: print("no bench line", e)
PY
done
