#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list + one full capture of the top kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"; cat $OUT/bench_reference.json
echo "== bench b200"; timeout 900 python bench.py > $OUT/bench_b200.json 2> $OUT/bench_b200.err; echo "rc=$?"; cat $OUT/bench_b200.json; tail -3 $OUT/bench_b200.err
echo "== bench b200 no flush"; timeout 900 python bench.py --no-flush --no-cpu-baseline > $OUT/bench_b200_noflush.json 2>> $OUT/bench_b200.err; echo "rc=$?"; cat $OUT/bench_b200_noflush.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full k1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_assoc -s 6 -c 3 -f -o $OUT/k1_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1; echo "rc=$?"
ls -la $OUT
