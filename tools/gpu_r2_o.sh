#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 1500 python -u -m pytest tests -m gpu --timeout 600 -x -q -p no:cacheprovider > gpurun_out/tests_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_all.log
timeout 300 python tools/pruned_probe.py > gpurun_out/pruned_probe.log 2>&1; cat gpurun_out/pruned_probe.log
timeout 300 python tools/pruned_profile.py 1 > gpurun_out/pruned_profile_b1.log 2>&1; grep -A70 "kernels per forward" gpurun_out/pruned_profile_b1.log | cut -c1-100
