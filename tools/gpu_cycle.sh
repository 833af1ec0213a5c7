#!/bin/bash
# One GPU cycle: gpu tests, default bench, optional extras selected by env (NCU_PAIR=1, SWEEP=1).  Logs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests -m gpu -q --timeout 120 -x -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -8 gpurun_out/tests.log
timeout 600 python -u bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
if [ -n "$NCU_PAIR" ]; then
  for M in $NCU_PAIR; do
    timeout 600 ncu --set full --clock-control none -k regex:gemm_tc -c 2 -o gpurun_out/prof_pair_$M -f python -u tools/cta2_probe.py --one $M > gpurun_out/ncu_pair_$M.log 2>&1; echo "ncu pair $M rc=$?"
  done
fi
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
