#!/bin/bash
TAG=${1:-r10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for P in 12500 25000 50000; do
for LPK in 1 2 4; do
  SRL_FAST_LPK=$LPK timeout 600 python bench.py --no-cpu-baseline --steps 10 --points $P > $OUT/bench_p${P}_lpk$LPK.json 2> $OUT/bench_p${P}_lpk$LPK.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_p${P}_lpk$LPK.json"))
    print("points=$P lpk=$LPK value %.1fM ms/step %.3f k1_avg_ms %.4f"%(d["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"]))
except Exception as e: print("failed", e)
PY
done; done
