#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'))"; }
for S in 5 8 12 16 32; do
CUM_STREAM_SMALL_STREAMS=64 CUM_STREAM_SMALL_MMAC=1000 timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk6.err | show "S=$S all small-M rm4"
CUM_STREAM_SMALL_STREAMS=64 CUM_STREAM_SMALL_MMAC=16 timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk6.err | show "S=$S small-M<=16MMAC rm4"
done
CUM_STREAM_SMALL_STREAMS=64 CUM_STREAM_SMALL_MMAC=1000 CUM_GEMM_SKINNY_RM=8 timeout 300 python bench.py --mode stream --model e6 --streams 8 --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk6.err | show "S=8 all small-M rm8"
timeout 300 python bench.py --mode stream --model e6 --streams 8 --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk6.err | show "S=8 default"
tail -n 3 gpurun_out/sk6.err
