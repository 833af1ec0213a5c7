#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 600 python -u -m pytest tests/test_gpu_ops.py -m gpu -rP --timeout 120 -x -q -p no:cacheprovider -k "fused or generic or shutdown" > gpurun_out/tests_fused.log 2>&1; echo "pytest fused rc=$?"
tail -15 gpurun_out/tests_fused.log
timeout 300 python tools/grad_probe.py 1.0 > gpurun_out/grad_probe.log 2>&1; tail -40 gpurun_out/grad_probe.log
timeout 900 python -u -m pytest tests/test_gpu_model.py tests/test_gpu_stream.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_model.log 2>&1; echo "pytest model rc=$?"
grep -E "^\[|passed|failed|Error|error" gpurun_out/tests_model.log | tail -30
for f in 1 0; do CUM_FUSED_ENDS=$f timeout 300 python bench.py --no-variants --no-cpu-baseline > gpurun_out/bench_fused_$f.json 2> gpurun_out/bench_fused_$f.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_fused_$f.json').read().strip().splitlines()[-1])
print('fused=$f', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
PY
done
