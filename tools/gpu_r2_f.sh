#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
for l in fused pytorch; do timeout 300 python bench.py --mode train --loss $l --steps 5 > gpurun_out/bench_train_$l.json 2> gpurun_out/bench_train_$l.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_train_$l.json').read().strip().splitlines()[-1])
print('train loss=$l', d['value'], d['ms_per_step'], d['config']['our_kernels_ms_per_step'], d['config']['final_loss'])
PY
done
