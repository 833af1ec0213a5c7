#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
run() { timeout 300 python bench.py --steps 10 --no-variants --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if 'gemm' in k})"; }
run base
CUM_HL16_CK=256 run ck256
CUM_HL16_CK=256 CUM_HL16_PK=128 run ck256_pk128
run base
CUM_HL16_CK=256 run ck256
CUM_HL16_CK=128 CUM_HL16_PK=128 run ck128_pk128
