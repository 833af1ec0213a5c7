#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_grad.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_grad.log 2>&1; echo "pytest grad rc=$?"
grep -E "^\[|passed|failed|^E  |Error" gpurun_out/tests_grad.log | tail -14
for f in 1 2; do timeout 300 python bench.py --mode train --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], d['config']['final_loss'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"
done
