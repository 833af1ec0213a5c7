#!/bin/bash
# gradient tests + training-step bench with CTA pairs on / off
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_grad.py -m gpu -q --timeout 180 -x -p no:cacheprovider > gpurun_out/tests_grad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_grad.log
tail -5 gpurun_out/tests_grad.log
for V in 1 0; do
  CUM_GEMM_CTA2=$V timeout 600 python -u bench.py --mode train --steps 3 --warmup 3 --math tf32x3 > gpurun_out/bench_train_cta2_$V.json 2> gpurun_out/bench_train_cta2_$V.err; echo "train CTA2=$V rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train_cta2_$V.json') if l.startswith('{')][-1])
print('CTA2=$V', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
PY
done
