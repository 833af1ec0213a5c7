#!/bin/bash
# round 2: 1-hop streaming A/Bs on one box -- CTA pairs off, scan-step channels per thread, fp16 state eager / graph
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['real_time_factor_per_stream'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('selective_scan','gemm','gemm_tap2','dwconv_silu','stream_shift')})"; }
run() { label=$1; shift; timeout 300 env "$@" python bench.py --mode stream --model e6 --streams 4096 --hops 1 --steps 30 --warmup 5 $EXTRA 2>>gpurun_out/u.err | show "$label"; }
EXTRA=""
run base X=1
run nopairs CUM_GEMM_CTA2=0
run cpt1 CUM_SCAN_STEP_CPT=1
run cpt4 CUM_SCAN_STEP_CPT=4
run nopdl CUM_PDL=0
EXTRA="--state-f16"
run f16_cpt4 X=1
run f16_cpt2 CUM_SCAN_STEP_CPT=2
run f16_cpt1 CUM_SCAN_STEP_CPT=1
EXTRA="--state-f16 --graph"
run f16_graph X=1
EXTRA="--hops 2"
EXTRA=""
timeout 300 python bench.py --mode stream --model e6 --streams 4096 --hops 2 --steps 20 --warmup 5 2>>gpurun_out/u.err | show h2
timeout 300 python bench.py --mode stream --model e6 --streams 2048 --hops 1 --steps 30 --warmup 5 2>>gpurun_out/u.err | show s2048_h1
timeout 300 python bench.py --mode stream --model e6 --streams 512 --hops 1 --steps 30 --warmup 5 2>>gpurun_out/u.err | show s512_h1
timeout 300 python bench.py --mode stream --model e6 --streams 512 --hops 1 --steps 30 --warmup 5 --graph 2>>gpurun_out/u.err | show s512_h1_graph
tail -n 3 gpurun_out/u.err
