#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
for i in 1 2; do timeout 300 python bench.py --steps 10 --no-variants --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('offline', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"
done
timeout 300 python tools/pruned_probe.py 2>&1 | grep "B=64"
