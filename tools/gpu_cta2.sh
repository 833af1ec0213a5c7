#!/bin/bash
# CTA-pair GEMM bring-up on the GPU box: probe (bounded), op tests, full gpu tests, A/B bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 300 python -u tools/cta2_probe.py > gpurun_out/cta2_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/cta2_probe.log
tail -45 gpurun_out/cta2_probe.log
if grep -q "probe rc=0" gpurun_out/cta2_probe.log; then
  timeout 900 python -u -m pytest tests -m gpu -q --timeout 120 -x -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
  tail -15 gpurun_out/tests.log
  timeout 600 python -u bench.py --steps 5 --warmup 3 > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench pair rc=$?"
  tail -c 1500 gpurun_out/bench_pair.json; tail -3 gpurun_out/bench_pair.err
  CUM_GEMM_CTA2=0 timeout 600 python -u bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single rc=$?"
  tail -c 1500 gpurun_out/bench_single.json; tail -3 gpurun_out/bench_single.err
fi
