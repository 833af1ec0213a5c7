#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
CUM_GEMM_CTA2_N128=1 timeout 900 python -u -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu --timeout 300 -q -x -p no:cacheprovider -k "gemm or golden or full_size" > gpurun_out/tests_n128.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tests_n128.log
for f in 1 0 1 0; do CUM_GEMM_CTA2_N128=$f timeout 300 python bench.py --no-variants --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n128 pairs=$f offline', d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if 'gemm' in k})"; done
