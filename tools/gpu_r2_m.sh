#!/bin/bash
# grad tests + train breakdown (A/B of the fused ReLU backward)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_grad.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_grad.log 2>&1; echo "pytest grad rc=$?"
grep -E "^\[|passed|failed|^E  |Error" gpurun_out/tests_grad.log | tail -14
for f in 1 0 1 0; do CUM_TRAIN_FUSED_RELU_BWD=$f timeout 300 python bench.py --mode train --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('train fused_relu_bwd=$f', d['value'], d['ms_per_step'], d['config']['final_loss'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('gemm','gemm_tap2','relu_bwd','colsum','glu_bwd','wgrad','selective_scan_bwd','selective_scan')})"
done
