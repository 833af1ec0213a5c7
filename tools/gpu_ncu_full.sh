#!/bin/bash
# One `ncu --set full` capture of the dominant kernels (1 GPU).  Reports land in gpurun_out/ (read back with ncu -i).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MATH=${MATH:-tf32x3}
# 3 warm-up steps + first timed step = 4 forwards before the one we capture; 44 GEMM launches per forward
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $((44*4)) -c 16 -f -o gpurun_out/prof_gemm_$MATH \
    python -u bench.py --steps 2 --warmup 3 --math $MATH --batch ${NCU_BATCH:-16} --no-cpu-baseline > gpurun_out/ncu_full_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selective_scan -s 12 -c 1 -f -o gpurun_out/prof_scan \
    python -u bench.py --steps 2 --warmup 3 --math $MATH --batch ${NCU_BATCH:-16} --no-cpu-baseline > gpurun_out/ncu_full_scan.log 2>&1; echo "ncu scan rc=$?"
ls -la gpurun_out/*.ncu-rep
