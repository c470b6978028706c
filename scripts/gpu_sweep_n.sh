#!/bin/bash
# k1 time vs keypoint count (wave quantisation check)
OUT=gpurun_out/${1:-nsweep}
mkdir -p $OUT
for n in 60000 90000 94000 100000 120000 189000; do
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --points $n > $OUT/b_$n.json 2> $OUT/b_$n.err
  python - <<PY
import json
d=json.load(open("$OUT/b_$n.json")); print($n, "value %.1fM k1 %.4f ms"%(d["value"]/1e6, d["roofline"]["k1_avg_ms"]))
PY
done
