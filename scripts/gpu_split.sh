#!/bin/bash
# split form (k1_scan + k1_fit): parity tests, then A/B timing against k1_fast
OUT=gpurun_out/${1:-split1}
mkdir -p $OUT
echo "== split tests"; timeout 900 python -m pytest tests -q -m gpu -k "split or hands_ambiguous or randomized" > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 20 $EXTRA > $OUT/b_$name.json 2> $OUT/b_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/b_$name.json")); print("$name", "value %.1fM e2e %.1fM k1 %.4f ms"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["k1_avg_ms"]))
except Exception as e: print("$name failed", e)
PY
}
EXTRA="--k1-variant 1" run fast X=1
EXTRA="--k1-variant 3" run split_l4_s8_f5 X=1
EXTRA="--k1-variant 3" run split_l4_s8_f4 SRL_FIT_MINB=4
EXTRA="--k1-variant 3" run split_l4_s8_f6 SRL_FIT_MINB=6
EXTRA="--k1-variant 3" run split_l2_s8_f5 SRL_SPLIT_LPK=2
bash scripts/gpu_split_ncu.sh ${1:-split1}_ncu 3 > $OUT/ncu.log 2>&1; grep "k1_\|k_" $OUT/ncu.log | grep avg
