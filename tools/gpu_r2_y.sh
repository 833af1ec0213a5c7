#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_ops.py tests/test_gpu_stream_tm.py tests/test_gpu_stream.py -m gpu --timeout 300 -x -q -p no:cacheprovider -k "scan or stream or tm or time_major or step" > gpurun_out/tests_y.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_y.log | tail -20
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['real_time_factor_per_stream'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('selective_scan','gemm','gemm_tap2','dwconv_silu','stream_shift')})"; }
run() { label=$1; shift; timeout 300 env "$@" python bench.py --mode stream --model e6 --streams 4096 --steps 20 --warmup 5 $EXTRA 2>>gpurun_out/y.err | show "$label"; }
for h in 1 2 4 8 16; do
EXTRA="--hops $h"
run h${h}_bulk X=1
run h${h}_old CUM_SCAN_STEP_BULK=0
done
EXTRA="--hops 4 --graph"
run h4_graph X=1
EXTRA="--hops 4 --state-f16"
run h4_f16 X=1
EXTRA="--hops 64"
run h64 X=1
tail -n 3 gpurun_out/y.err
