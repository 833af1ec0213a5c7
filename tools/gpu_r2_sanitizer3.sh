#!/bin/bash
# compute-sanitizer over the small-M GEMM path and both streaming sessions at 1-4 streams (after profiles/r02_sanitizer.md part 2)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_stream_tm.py tests/test_gpu_stream.py -m gpu -q -x --timeout 1200 -p no:cacheprovider -k "small_m or module_feed or matches_stream_oracle or edge_cases or bit_identical or pruned" > gpurun_out/sanitizer3_mem.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer3_mem.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_stream_tm.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "small_m" > gpurun_out/sanitizer3_race.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer3_race.log | tail -3
