#!/bin/bash
# last GPU call of round 2: the tests added at the end of the round first (GPU vs the compiled reference, the link-time patched
# reference on the CUDA backend), then the whole GPU suite, smoke, the driver's bench line, ordering times, a launch list,
# the CPU arm.  Every step has its own timeout; results under gpurun_out/$1.
OUT=gpurun_out/${1:-f1}
mkdir -p $OUT
export PYTHONPATH=$PWD
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "compiled_reference or gpu_backend" > $OUT/new_tests.log 2>&1; el "new tests rc=$?"; tail -12 $OUT/new_tests.log
timeout 420 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; el "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; el "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; el "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n1.json")); c=d.get("cpu_baseline") or {}
    print("value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f frac %.3f clocks %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"].get("ms_per_step",0),d["roofline"]["k1_avg_ms"],d["roofline"]["frac"],d["clocks"]))
    print("cpu_baseline", {k:v for k,v in c.items() if k not in ("sample","library")})
except Exception as e: print("no bench line", e); print(open("$OUT/bench_n1.err").read()[-1500:])
PY
timeout 100 python scripts/order_time.py 100000 200 > $OUT/order_time.txt 2>&1; el "order_time rc=$?"; tail -6 $OUT/order_time.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --no-cpu-baseline --no-cfg5 --steps 2 --warmup 1 > $OUT/ncu_bench.log 2>&1; el "ncu rc=$?"
python - <<PY
import csv,collections
try:
    rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    agg=collections.defaultdict(list)
    for r in rows[1:]:
        try: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
        except: pass
    for k,v in agg.items(): print("%-62s n=%3d avg %.1f us"%(k,len(v),sum(v)/len(v)/1000))
except Exception as e: print("no launch list", e)
PY
timeout 240 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; el "reference arm rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_reference.json")); c=d["cpu_baseline"]
    print("reference arm value %.2fM ms/step %.1f"%(d["value"]/1e6,d["ms_per_step"]), {k:v for k,v in c.items() if k not in ("sample","library")}, d["config"].get("map_build_s"))
except Exception as e: print("no reference line", e); print(open("$OUT/bench_reference.err").read()[-1500:])
PY
