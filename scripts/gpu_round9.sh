#!/bin/bash
TAG=${1:-r09}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== IPC exchange test (2 processes, 1 GPU)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k fused_peer > $OUT/pytest_ipc.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_ipc.log
echo "== pytest -m gpu (rest)"; timeout 1500 python -m pytest tests -q -m gpu -k "not fused_peer" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
