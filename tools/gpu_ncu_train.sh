#!/bin/bash
# launch list (durations) of one training step + refreshed forward captures for the final kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 760 --csv --log-file gpurun_out/launches_train_f16x3.csv \
    python -u bench.py --mode train --steps 1 --warmup 3 --math f16x3 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
tail -2 gpurun_out/ncu_train.log | cut -c1-300
