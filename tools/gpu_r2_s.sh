#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -k regex:"gemm_tc_kernel|split_planes" -s 588 -c 196 --csv --log-file gpurun_out/ncu_train_gemms.csv \
    python -u bench.py --mode train --steps 1 --warmup 3 --math f16x3 > gpurun_out/ncu_train2.log 2>&1; echo "ncu rc=$?"
tail -1 gpurun_out/ncu_train2.log | cut -c1-200
