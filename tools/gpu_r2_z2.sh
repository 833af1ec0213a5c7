#!/bin/bash
# A/B of the tile-width rule: CUM_GEMM_FILL=0 (256 unless n <= 128), 1 (128 when the doubled tile count fits one wave), default (wave-quantised cost)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'), {k:v['ms_per_step'] for k,v in d.get('kernels',{}).items() if k in ('selective_scan','gemm','gemm_tap2')})"; }
run() { label=$1; shift; timeout 300 env "$@" python bench.py --mode stream --model e6 --steps 30 --warmup 5 $EXTRA 2>>gpurun_out/z2.err | show "$label"; }
for mode in 2 1 0; do
EXTRA="--streams 4096 --hops 1"; run h1_fill$mode CUM_GEMM_FILL=$mode
EXTRA="--streams 4096 --hops 1 --graph"; run h1_graph_fill$mode CUM_GEMM_FILL=$mode
EXTRA="--streams 4096 --hops 2 --graph"; run h2_graph_fill$mode CUM_GEMM_FILL=$mode
EXTRA="--streams 512 --hops 1 --graph"; run s512_fill$mode CUM_GEMM_FILL=$mode
EXTRA="--streams 1 --hops 1 --graph --steps 100"; run s1_fill$mode CUM_GEMM_FILL=$mode
done
for mode in 2 0; do CUM_GEMM_FILL=$mode timeout 300 python bench.py --mode sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('sweep fill$mode', [(p['batch'],p['clip_seconds'],p['ms_per_step']) for p in d['sweep']])"; done
tail -n 3 gpurun_out/z2.err
