#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 600 python -u -m pytest tests/test_gpu_ops.py -m gpu --timeout 300 -x -q -p no:cacheprovider -k "standalone" > gpurun_out/tests_q2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_q2.log | tail -12
