#!/bin/bash
TAG=${1:-r07}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench b200 default"; timeout 900 python bench.py > $OUT/bench_b200.json 2> $OUT/bench_b200.err; echo "rc=$?"; cat $OUT/bench_b200.json; tail -3 $OUT/bench_b200.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "rc=$?"
