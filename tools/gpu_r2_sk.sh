#!/bin/bash
# skinny (small-M CUDA-core) GEMM path: streaming / op tests, then few-stream latency in both layouts with the path on / off
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_stream.py tests/test_gpu_stream_tm.py tests/test_gpu_ops.py -m gpu --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_sk.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_sk.log | tail -12
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'), d['config'].get('buffer_layout'))"; }
for sk in 1 0; do for S in 1 2; do for lay in stream_major time_major; do
CUM_GEMM_SKINNY=$sk timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph --layout $lay 2>>gpurun_out/sk.err | show "skinny=$sk S=$S $lay graph"
done; done; done
CUM_GEMM_SKINNY=1 timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 1 --steps 100 --warmup 5 --layout stream_major 2>>gpurun_out/sk.err | show "skinny=1 S=1 eager"
CUM_GEMM_SKINNY=1 timeout 300 python bench.py --mode stream --model e8 --streams 1 --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk.err | show "skinny=1 E8 S=1 graph"
CUM_GEMM_SKINNY=0 timeout 300 python bench.py --mode stream --model e8 --streams 1 --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk.err | show "skinny=0 E8 S=1 graph"
tail -n 3 gpurun_out/sk.err
