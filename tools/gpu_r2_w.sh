#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['real_time_factor_per_stream'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('selective_scan','gemm','gemm_tap2','dwconv_silu','stream_shift')})"; }
run() { label=$1; shift; timeout 300 env "$@" python bench.py --mode stream --model e6 --streams 4096 --steps 30 --warmup 5 $EXTRA 2>>gpurun_out/w.err | show "$label"; }
EXTRA="--hops 1"
run h1_bulk X=1
run h1_old CUM_SCAN_STEP_BULK=0
EXTRA="--hops 1 --state-f16"
run h1_f16_bulk X=1
EXTRA="--hops 2"
run h2_bulk X=1
EXTRA="--hops 2 --state-f16"
run h2_f16_bulk X=1
EXTRA="--hops 1 --graph"
run h1_bulk_graph X=1
EXTRA="--hops 1 --graph --state-f16"
run h1_f16_bulk_graph X=1
tail -n 3 gpurun_out/w.err
