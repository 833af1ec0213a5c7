#!/bin/bash
# round-2 ncu evidence (1 GPU): launch list of one forward, --set full of the fused end blocks / scan, streaming state-update scan,
# training wgrad + reverse scan.  Reports land in gpurun_out/ (<= 64 MiB in total).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BENCH="python -u bench.py --no-cpu-baseline --no-variants --no-extras"
K='gemm_tc_kernel|fused_end_kernel|selective_scan|conv_in|convt_out|ln_residual|dwconv|wave_normalize'
timeout 900 ncu -k regex:"$K" \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_f16x3.csv \
    $BENCH --steps 1 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# 4 forwards before the captured one (3 warm-up + 1): fused_end_kernel launches twice per forward
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_end_kernel -s 8 -c 2 -f -o gpurun_out/prof_fused_ends \
    $BENCH --steps 2 --warmup 3 > gpurun_out/ncu_full_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:selective_scan_fwd -s 12 -c 1 -f -o gpurun_out/prof_scan_r02 \
    $BENCH --steps 2 --warmup 3 > gpurun_out/ncu_full_scan.log 2>&1; echo "ncu scan rc=$?"
# streaming, 4096 streams x 1 hop: the state-update scan (one launch per Mamba layer and step)
timeout 600 ncu --set full --clock-control none -k regex:selective_scan_step -s 30 -c 3 -f -o gpurun_out/prof_scan_step \
    python -u bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 5 > gpurun_out/ncu_full_step.log 2>&1; echo "ncu step rc=$?"
# training: weight-gradient GEMMs (+ their transposing pre-passes) and the reverse scan
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
    -k regex:"transpose_pitch|gemm_tc_kernel<1, 256, -3|gemm_tc_kernel<1, 128, -3|selective_scan_bwd" -s 400 -c 60 -f -o gpurun_out/prof_train_r02 \
    python -u bench.py --mode train --steps 2 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out; find gpurun_out -name "*.log" -size +1M -delete
