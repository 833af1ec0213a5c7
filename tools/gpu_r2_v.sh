#!/bin/bash
# round 2: TMA-staged state-update scan -- tests, then same-box A/B at 4096 streams x 1 / 2 hops, fp32 and fp16 state
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_ops.py tests/test_gpu_stream_tm.py tests/test_gpu_stream.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider -k "scan or stream or tm or time_major or step" > gpurun_out/tests_v.log 2>&1; echo "pytest rc=$?"
grep -E "^\[|passed|failed|^E  |Error" gpurun_out/tests_v.log | tail -20
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['real_time_factor_per_stream'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('selective_scan','gemm','gemm_tap2','dwconv_silu','stream_shift')})"; }
run() { label=$1; shift; timeout 300 env "$@" python bench.py --mode stream --model e6 --streams 4096 --steps 30 --warmup 5 $EXTRA 2>>gpurun_out/v.err | show "$label"; }
EXTRA="--hops 1"
run h1_bulk X=1
run h1_old CUM_SCAN_STEP_BULK=0
EXTRA="--hops 1 --state-f16"
run h1_f16_bulk X=1
run h1_f16_old CUM_SCAN_STEP_BULK=0
EXTRA="--hops 2"
run h2_bulk X=1
run h2_old CUM_SCAN_STEP_BULK=0
EXTRA="--hops 2 --state-f16"
run h2_f16_bulk X=1
EXTRA="--hops 1 --graph"
run h1_bulk_graph X=1
tail -n 3 gpurun_out/v.err
