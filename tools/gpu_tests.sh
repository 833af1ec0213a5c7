#!/bin/bash
# Run on the GPU box: the gpu-marked tests with per-test timeouts; everything logged under gpurun_out/ so a timeout still leaves evidence.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
nvidia-smi -L > gpurun_out/diag.log 2>&1
timeout ${TEST_TIMEOUT:-1200} python -u -m pytest tests -m gpu -rP --timeout 300 -q -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
grep -E "^\[|passed|failed|^E  |rc=" gpurun_out/tests.log | tail -60
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
