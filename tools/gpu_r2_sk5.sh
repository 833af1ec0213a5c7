#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_stream.py tests/test_gpu_stream_tm.py -m gpu --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_sk5.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_sk5.log | tail -6
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'))"; }
for S in 1 2 4; do
timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 200 --warmup 5 --graph 2>>gpurun_out/sk5.err | show "S=$S"
done
timeout 300 python bench.py --mode stream --model e8 --streams 1 --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk5.err | show "E8 S=1"
timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 1 --steps 200 --warmup 5 --graph --layout stream_major 2>>gpurun_out/sk5.err | show "S=1 stream_major"
tail -n 3 gpurun_out/sk5.err
