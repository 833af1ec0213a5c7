#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_grad.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_grad.log 2>&1; echo "pytest grad rc=$?"
grep -E "^\[|passed|failed|^E  |Error" gpurun_out/tests_grad.log | tail -30
for f in 1 0; do CUM_TRAIN_F16_BWD=$f timeout 300 python bench.py --mode train --steps 5 > gpurun_out/bench_train_f16bwd$f.json 2> gpurun_out/bench_train_f16bwd$f.err; tail -2 gpurun_out/bench_train_f16bwd$f.err | cut -c1-200; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_train_f16bwd$f.json').read().strip().splitlines()[-1])
print('train f16_bwd=$f', d['value'], d['ms_per_step'], d['config']['final_loss'], {k:(v['ms_per_step'],v['tflops']) for k,v in list(d['kernels'].items())[:8]})
PY
done
