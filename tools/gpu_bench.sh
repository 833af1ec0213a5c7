#!/bin/bash
# Run on the GPU box: bench lines (both math modes) + an ncu launch list.  Output under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for MATH in ${MATHS:-fp32 tf32x3}; do
  timeout 600 python -u bench.py --steps ${STEPS:-5} --warmup 3 --math $MATH ${BENCH_ARGS} > gpurun_out/bench_$MATH.json 2> gpurun_out/bench_$MATH.err; echo "bench $MATH rc=$?"
  tail -c 3000 gpurun_out/bench_$MATH.json; tail -5 gpurun_out/bench_$MATH.err
done
if [ -n "$NCU_MATH" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-700} --csv --log-file gpurun_out/launches_$NCU_MATH.csv \
      python -u bench.py --steps 1 --warmup 3 --math $NCU_MATH --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_bench_$NCU_MATH.log 2>&1; echo "ncu rc=$?"
  tail -3 gpurun_out/ncu_bench_$NCU_MATH.log
fi
