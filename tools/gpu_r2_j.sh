#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 1200 python -u -m pytest tests -m gpu --timeout 300 -q -x -p no:cacheprovider > gpurun_out/tests_all.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tests_all.log
timeout 300 python bench.py --no-variants --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('offline', d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"
