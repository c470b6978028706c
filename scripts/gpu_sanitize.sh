#!/bin/bash
# compute-sanitizer passes over the small diagnostic (map build, insert, scan-matching passes, ESIKF update)
TAG=${1:-san1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== new randomized parity test"; timeout 900 python -m pytest tests -q -m gpu -k "randomized" > $OUT/pytest_rand.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_rand.log
for tool in racecheck synccheck initcheck; do
  echo "== $tool"; timeout 700 compute-sanitizer --tool $tool python scripts/gpu_diag.py small > $OUT/$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/$tool.log | head -8
done
