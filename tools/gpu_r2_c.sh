#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 600 python -u -m pytest tests/test_gpu_ops.py -m gpu -rP --timeout 120 -x -q -p no:cacheprovider -k "fused" > gpurun_out/tests_fused.log 2>&1; echo "pytest fused rc=$?"
tail -5 gpurun_out/tests_fused.log
timeout 900 python -u -m pytest tests/test_gpu_grad.py tests/test_gpu_stream.py tests/test_gpu_model.py -m gpu -rP --timeout 300 -q -p no:cacheprovider -k "e8_full_grad or accumulate or weight_updates or golden" > gpurun_out/tests_model.log 2>&1; echo "pytest model rc=$?"
grep -E "^\[|passed|failed|^E  " gpurun_out/tests_model.log | tail -30
for f in 1 0; do CUM_FUSED_ENDS=$f timeout 300 python bench.py --no-variants --no-cpu-baseline --no-extras > gpurun_out/bench_fused_$f.json 2> gpurun_out/bench_fused_$f.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_fused_$f.json').read().strip().splitlines()[-1])
print('fused=$f', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
PY
done
