#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_stream.py tests/test_gpu_stream_tm.py tests/test_gpu_ops.py -m gpu --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_sk3.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_sk3.log | tail -8
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'))"; }
for S in 1 2 4 8 16 64 256; do for MB in 16 0; do
CUM_GEMM_SKINNY=$([ $MB = 0 ] && echo 0 || echo 1) timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk3.err | show "S=$S skinny=$MB"
done; done
for MB in 3 13; do for S in 2 4 8; do
CUM_STREAM_SMALL_MMAC=$MB timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk3.err | show "S=$S budget_MMAC=$MB"
done; done
tail -n 3 gpurun_out/sk3.err
