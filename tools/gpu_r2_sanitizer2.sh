#!/bin/bash
# compute-sanitizer over the code added after profiles/r02_sanitizer.md: time-major session, plane-major GEMM operands, strided kernels,
# stream_shift, TMA-staged state-update scan (memcheck: all of tests/test_gpu_stream_tm.py + the scan / stream tests; racecheck: the
# stream_shift / strided / step-scan kernels)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_stream_tm.py tests/test_gpu_stream.py -m gpu -q -x --timeout 1200 -p no:cacheprovider > gpurun_out/sanitizer2_tm.log 2>&1; echo "memcheck tm rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer2_tm.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "selective_scan or gemm_as_strided or pointwise" > gpurun_out/sanitizer2_ops.log 2>&1; echo "memcheck ops rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer2_ops.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_stream_tm.py tests/test_gpu_ops.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "stream_shift or strided_wave or step_scan or selective_scan_matches" > gpurun_out/sanitizer2_race.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer2_race.log | tail -3
