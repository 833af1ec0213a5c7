#!/bin/bash
# Run on the GPU box: staged diagnostic, then the gpu-marked tests with per-test timeouts; everything logged
# unbuffered under gpurun_out/ so a timeout still leaves evidence.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
nvidia-smi -L > gpurun_out/diag.log 2>&1
timeout 300 python -u tools/gpu_diag.py "$@" >> gpurun_out/diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/diag.log
tail -25 gpurun_out/diag.log
timeout ${TEST_TIMEOUT:-600} python -u -m pytest tests -m gpu -v -rP --timeout 120 -x -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -40 gpurun_out/tests.log
