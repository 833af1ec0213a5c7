#!/bin/bash
# re-capture of the offline-forward launch list on the final round-2 code (profiles/r02_launches_f16x3.{csv,md}, r02_gemm_traffic_f16x3.json)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BENCH="python -u bench.py --no-cpu-baseline --no-variants --no-extras"
K='gemm_tc_kernel|fused_end_kernel|selective_scan|conv_in|convt_out|ln_residual|dwconv|wave_normalize'
timeout 900 ncu -k regex:"$K" \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_f16x3.csv \
    $BENCH --steps 1 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/launches_r02_f16x3.csv
