#!/bin/bash
# A/B of the tile-width rule for under-filled GEMMs (CUM_GEMM_FILL=0: round-2 rule), streaming 1 hop / pruned forward / single stream
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -u -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu --timeout 300 -x -q -p no:cacheprovider -k "gemm or pruned or golden or oracle" > gpurun_out/tests_z.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_z.log | tail -5
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'), {k:v['ms_per_step'] for k,v in d.get('kernels',{}).items() if k in ('selective_scan','gemm','gemm_tap2')})"; }
run() { label=$1; shift; timeout 300 env "$@" python bench.py --mode stream --model e6 --steps 30 --warmup 5 $EXTRA 2>>gpurun_out/z.err | show "$label"; }
EXTRA="--streams 4096 --hops 1"
run h1_fill X=1
run h1_nofill CUM_GEMM_FILL=0
EXTRA="--streams 4096 --hops 1 --graph"
run h1_fill_graph X=1
run h1_nofill_graph CUM_GEMM_FILL=0
EXTRA="--streams 1 --hops 1 --graph --steps 100"
run s1_fill X=1
run s1_nofill CUM_GEMM_FILL=0
EXTRA="--streams 512 --hops 1 --graph"
run s512_fill X=1
run s512_nofill CUM_GEMM_FILL=0
for e in 1 0; do CUM_GEMM_FILL=$e timeout 300 python tools/pruned_probe.py 2>&1 | tail -4; done
tail -n 3 gpurun_out/z.err
