#!/bin/bash
# the single-launch sweep ordering: parity subset, then e2e A/B against the CUB order, then its launch time
OUT=gpurun_out/${1:-o1}
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x -k "split or randomized or pass_matches or golden or deterministic or iekf or ragged or full_size or optimize_host" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for v in 1 0 1 0; do
SRL_CLUSTER_ORDER=$v timeout 300 python bench.py --no-cpu-baseline --steps 40 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; tail -2 $OUT/bench_$v.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$v.json")); print("cluster_order=$v value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f | %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["config"]["sweep_order"][:40]))
except Exception as e: print("no bench line", e)
PY
done
SRL_DEVICE_LOOP=0 bash scripts/gpu_ncu_list.sh ${1:-o1}_l 60 | grep "sweep\|Radix\|mismatch"
