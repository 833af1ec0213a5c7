#!/bin/bash
# round-2 ncu evidence for the time-major streaming step (4096 streams x 1 hop): --set full of the TMA-staged state-update scan, and the
# launch list of one steady-state call
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selective_scan_step_bulk -s 30 -c 3 -f -o gpurun_out/prof_scan_step_bulk \
    python -u bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 5 > gpurun_out/ncu_full_step_bulk.log 2>&1; echo "ncu step bulk rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:selective_scan_step_bulk -s 30 -c 3 -f -o gpurun_out/prof_scan_step_bulk_f16 \
    python -u bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 5 --state-f16 > gpurun_out/ncu_full_step_bulk_f16.log 2>&1; echo "ncu step bulk f16 rc=$?"
# launch list: 8 calls before the captured one (1 prime + 5 warm-up + ...): capture the kernels of the LAST timed call region broadly, post-filter here
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"cum::" -c 2000 --csv --log-file gpurun_out/launches_stream_h1.csv \
    python -u bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 2 --warmup 3 > gpurun_out/ncu_launches_stream.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_stream_h1.csv; du -sh gpurun_out
