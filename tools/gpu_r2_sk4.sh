#!/bin/bash
# final small-M rule (sessions of <= 4 streams; automatic: <= 4 output rows): full suite, then stream-count sweep with the path on / off
bash tools/gpu_tests.sh
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'))"; }
for S in 1 2 4 5 8 64; do for SK in 1 0; do
CUM_GEMM_SKINNY=$SK timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk4.err | show "S=$S small_m=$SK"
done; done
CUM_GEMM_SKINNY=1 timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 4 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk4.err | show "S=1 h4 small_m=1"
CUM_GEMM_SKINNY=0 timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 4 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk4.err | show "S=1 h4 small_m=0"
tail -n 3 gpurun_out/sk4.err
