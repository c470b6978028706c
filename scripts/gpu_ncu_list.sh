#!/bin/bash
# per-kernel launch times of a short bench run (cold-cache, serialised: shares, not absolutes)
OUT=gpurun_out/${1:-l1}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${2:-80} --csv --log-file $OUT/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $OUT/b.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items():
    print("%-72s n=%3d avg %.1f us  min %.1f max %.1f"%(k,len(v),sum(v)/len(v)/1000,min(v)/1000,max(v)/1000))
PY
