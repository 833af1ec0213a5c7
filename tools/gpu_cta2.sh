#!/bin/bash
# CTA-pair GEMM A/B on the GPU box: probe (bounded), op tests, bench with pairs forced on / off.  Logs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 300 python -u tools/cta2_probe.py > gpurun_out/cta2_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/cta2_probe.log
grep -v "= 0.0$" gpurun_out/cta2_probe.log | tail -30
if grep -q "probe rc=0" gpurun_out/cta2_probe.log; then
  timeout 900 python -u -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -x -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
  tail -5 gpurun_out/tests.log
  for V in 1 0; do
    CUM_GEMM_CTA2=$V timeout 600 python -u bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cta2_$V.json 2> gpurun_out/bench_cta2_$V.err; echo "bench CTA2=$V rc=$?"
    python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_cta2_$V.json') if l.startswith('{')][-1])
print('CTA2=$V', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, {k:v['ms_per_step'] for k,v in d['variants'].items()})
PY
  done
fi
