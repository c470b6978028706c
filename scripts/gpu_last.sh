#!/bin/bash
# last check of the round: full GPU test suite, smoke, the default bench line twice
OUT=gpurun_out/${1:-last}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
for i in 1 2; do
timeout 600 python bench.py > $OUT/bench_b200_$i.json 2> $OUT/bench_b200_$i.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_b200_$i.json")); print("value %.1fM e2e %.1fM ms/step %.3f k1 %.4f frac %.3f stats %s stream %.0f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["frac"],d["ms_per_step_stats"],d["streaming"]["sweeps_per_s"]))
PY
done
