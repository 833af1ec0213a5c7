#!/bin/bash
# A/B of an environment switch on the GPU box: gpu tests once, then the default bench for every value of $AB_VAR in $AB_VALUES.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests -m gpu -q --timeout 120 -x -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -12 gpurun_out/tests.log
for V in ${AB_VALUES:-1 0}; do
  env ${AB_VAR:-CUM_HL16}=$V timeout 600 python -u bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_ab_$V.json 2> gpurun_out/bench_ab_$V.err; echo "bench ${AB_VAR:-CUM_HL16}=$V rc=$?"; tail -3 gpurun_out/bench_ab_$V.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_ab_$V.json') if l.startswith('{')][-1])
    print('${AB_VAR:-CUM_HL16}=$V', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, {k:v['ms_per_step'] for k,v in d.get('variants',{}).items()})
except Exception as e: print('no bench line', e)
PY
done
