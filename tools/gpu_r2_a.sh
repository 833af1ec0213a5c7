#!/bin/bash
# round 2, call A: full GPU test suite (incl. the new parity-gap tests) + the default bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
nvidia-smi -L > gpurun_out/diag.log 2>&1
timeout ${TEST_TIMEOUT:-900} python -u -m pytest tests -m gpu -rP --timeout 300 -x -q -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
grep -E "^\[|max-abs|worst|passed|failed|rc=" gpurun_out/tests.log | tail -40
timeout 600 python bench.py --no-variants > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2a.json
