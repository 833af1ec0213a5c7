#!/bin/bash
# ncu captures of the dominant kernels (1 GPU).  Reports land in gpurun_out/ (<= 64 MiB in total!), read back with ncu -i.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MATH=${MATH:-f16x3}
B=${NCU_BATCH:-16}
BENCH="python -u bench.py --steps 2 --warmup 3 --math $MATH --no-cpu-baseline --no-variants"
# 3 warm-up steps + first timed step = 4 forwards before the one we capture; 44 GEMM launches per forward
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc -s $((44*4)) -c 8 -f -o gpurun_out/prof_gemm_$MATH \
    $BENCH --batch $B > gpurun_out/ncu_full_gemm.log 2>&1; echo "ncu gemm full rc=$?"
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none -k regex:gemm_tc -s $((44*4)) -c 44 -f -o gpurun_out/prof_gemm_all44_$MATH \
    $BENCH --batch 64 > gpurun_out/ncu_all44.log 2>&1; echo "ncu gemm all44 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selective_scan -s 12 -c 1 -f -o gpurun_out/prof_scan \
    $BENCH --batch $B > gpurun_out/ncu_full_scan.log 2>&1; echo "ncu scan rc=$?"
# launch list of one whole forward (our kernels only): time, DRAM bytes, tensor-pipe activity per launch -> tools/launch_table.py
timeout 900 ncu -k regex:"gemm_tc_kernel|selective_scan|conv_in|convt_out|ln_residual|dwconv|wave_normalize" \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 500 --csv --log-file gpurun_out/launches_$MATH.csv \
    python -u bench.py --steps 1 --warmup 3 --math $MATH --no-cpu-baseline --no-variants > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out; find gpurun_out -name "*.log" -size +1M -delete
