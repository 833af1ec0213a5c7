#!/bin/bash
# round 2, session 3: grad tests with the ReLU backward fused into the dgrad epilogue + same-box A/B + ncu launch list of one training step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_grad.py tests/test_gpu_ops.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider -k "grad or gemm" > gpurun_out/tests_grad.log 2>&1; echo "pytest grad rc=$?"
grep -E "^\[|passed|failed|^E  |Error" gpurun_out/tests_grad.log | tail -14
for f in 1 0 1 0; do CUM_TRAIN_FUSED_RELU_BWD=$f timeout 300 python bench.py --mode train --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('train fused_relu_bwd=$f', d['value'], d['ms_per_step'], d['config']['final_loss'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('gemm','gemm_tap2','relu_bwd','colsum','glu_bwd','wgrad','selective_scan_bwd')})"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none -k regex:"scan_bwd|rowblock|gemm_tc_kernel.*Li1EEE|gemm_tc_kernel.*Li2EEE|selective_scan_fwd" -c 120 --csv --log-file gpurun_out/ncu_train_kernels.csv \
    python -u bench.py --mode train --steps 1 --warmup 1 --math f16x3 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
tail -2 gpurun_out/ncu_train.log | cut -c1-300
