#!/bin/bash
# round 2: device-resident loop check: iekf tests, whole GPU suite, bench with the device loop and with the host loop, launch list
OUT=gpurun_out/${1:-r2a}
mkdir -p $OUT
(ncu --target-processes all printenv 2>/dev/null | grep -i "NV\|CUDA\|inject" > $OUT/ncu_env.txt) || true
timeout 600 python -m pytest tests -q -m gpu -x -k "iekf or device_resident or optimize_host or streaming" > $OUT/pytest_iekf.log 2>&1; echo "pytest iekf rc=$?"; tail -5 $OUT/pytest_iekf.log
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -15 $OUT/pytest_gpu.log
for dl in 1 0; do
SRL_DEVICE_LOOP=$dl timeout 600 python bench.py --no-cpu-baseline --steps 30 > $OUT/bench_dl$dl.json 2> $OUT/bench_dl$dl.err; echo "bench dl=$dl rc=$?"; tail -3 $OUT/bench_dl$dl.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_dl$dl.json")); print("dl=$dl value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f launches %d step %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"],d["gpu_launches"],d.get("iekf_step")))
except Exception as e: print("no bench line", e)
PY
done
bash scripts/gpu_ncu_list.sh ${1:-r2a}_l 100
