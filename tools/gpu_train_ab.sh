#!/bin/bash
# gradient tests + training-step bench A/B of an environment switch (default: f16x3 forward GEMMs on / off)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_grad.py -m gpu -q --timeout 180 -x -p no:cacheprovider -rP > gpurun_out/tests_grad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_grad.log
grep "worst per-tensor\|passed\|failed\|rc=" gpurun_out/tests_grad.log | tail -12
for V in ${AB_VALUES:-1 0}; do
  env ${AB_VAR:-CUM_TRAIN_F16_FWD}=$V timeout 600 python -u bench.py --mode train --steps 3 --warmup 3 ${TRAIN_ARGS} > gpurun_out/bench_train_ab_$V.json 2> gpurun_out/bench_train_ab_$V.err; echo "train ${AB_VAR:-CUM_TRAIN_F16_FWD}=$V rc=$?"; tail -2 gpurun_out/bench_train_ab_$V.err | cut -c1-300
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train_ab_$V.json') if l.startswith('{')][-1])
print('$V', d['value'], d['ms_per_step'], d['config']['final_loss'], {k:v['ms_per_step'] for k,v in list(d['kernels'].items())[:8]})
PY
done
