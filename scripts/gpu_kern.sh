#!/bin/bash
# kernel change check: parity subset, bench (device loop), per-kernel launch times (ncu, host loop)
OUT=gpurun_out/${1:-k1}
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x -k "${2:-split or hands_ambiguous or randomized or pass_matches or golden or host_plane}" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 30 > $OUT/bench$i.json 2> $OUT/bench$i.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench$i.json")); print("value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f step_cycles %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"],d["iekf_step"]["sm_cycles_sums_to_pose"]))
except Exception as e: print("no bench line", e)
PY
done
bash scripts/gpu_ncu_list.sh ${1:-k1}_l 70 | grep "k1_\|k_sweep\|Radix"
