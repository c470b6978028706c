#!/bin/bash
OUT=gpurun_out/${1:-fs}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 $OUT/pytest_gpu.log
export SRL_DEVICE_LOOP=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_fit -s 4 -c 1 -o /tmp/fit -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $OUT/b.log 2>&1
ncu -i /tmp/fit.ncu-rep --page source --csv > $OUT/k1_fit_source.csv 2>/dev/null
ncu -i /tmp/fit.ncu-rep --page details > $OUT/k1_fit_details.txt 2>&1
grep -E "^\s+Duration|Executed Ipc Active|Registers Per|Achieved Occupancy" $OUT/k1_fit_details.txt
