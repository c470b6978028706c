#!/bin/bash
# kernel change check: parity subset, bench at 100k and at 12.5k keypoints (an 8-GPU shard), per-kernel launch times (ncu, host loop)
OUT=gpurun_out/${1:-k1}
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x -k "${2:-split or hands_ambiguous or randomized or pass_matches or golden or host_plane or deterministic or iekf or ragged}" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for P in 100000 100000 12500; do
timeout 300 python bench.py --no-cpu-baseline --steps 40 --points $P > $OUT/bench_$P.json 2> $OUT/bench_$P.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$P.json")); print("P=$P value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f timleg %.3f step_cycles %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"],d["roofline"]["timing_leg_ms_per_step"],d["iekf_step"]["sm_cycles_sums_to_pose"]))
except Exception as e: print("no bench line", e)
PY
done
if [ -z "$3" ]; then SRL_DEVICE_LOOP=0 bash scripts/gpu_ncu_list.sh ${1:-k1}_l 70 | grep "k1_\|k_sweep\|Radix"; fi
