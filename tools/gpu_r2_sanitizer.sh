#!/bin/bash
# compute-sanitizer over the GPU tests (memcheck: every kernel; racecheck: the shared-memory kernels outside the TMA / mbarrier code)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py tests/test_gpu_stream.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "not 160000 and not 80126" > gpurun_out/sanitizer_ops.log 2>&1; echo "memcheck ops rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_ops.log | tail -3
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_grad.py tests/test_gpu_model.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "not vs_exact_fp32 and not benchmark_clip" > gpurun_out/sanitizer_model.log 2>&1; echo "memcheck model rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_model.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "selective_scan or wave_ends or conv_in or causal_conv or layer_norm or importance" > gpurun_out/sanitizer_race.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_race.log | tail -3
