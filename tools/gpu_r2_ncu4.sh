#!/bin/bash
# ncu evidence for the single-stream call (E6 full, 1 stream x 1 hop): launch list of one steady-state call
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"_kernel" -c 3000 --csv --log-file gpurun_out/launches_stream_s1_h1.csv \
    python -u bench.py --mode stream --model e6 --streams 1 --hops 1 --steps 2 --warmup 6 > gpurun_out/ncu_launches_stream_s1.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/launches_stream_s1_h1.csv
