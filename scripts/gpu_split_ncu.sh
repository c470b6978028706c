#!/bin/bash
OUT=gpurun_out/${1:-split_ncu}
mkdir -p $OUT
V=${2:-3}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 --k1-variant $V > $OUT/b.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items(): print("%-62s n=%3d avg %.1f us"%(k,len(v),sum(v)/len(v)/1000))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k1_scan|k1_fit" -s 6 -c 2 -o $OUT/split_full python bench.py --no-cpu-baseline --steps 2 --warmup 1 --k1-variant $V > $OUT/b2.log 2>&1
ls -la $OUT
