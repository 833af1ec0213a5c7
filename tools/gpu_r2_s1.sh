#!/bin/bash
# few streams: stream-major vs time-major session from the graph (and eager), E6 full
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'), d['config'].get('buffer_layout'), d['gpu_launches'])"; }
for S in 1 4 16 32 64; do for lay in stream_major time_major; do
timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph --layout $lay 2>>gpurun_out/s1.err | show "S=$S $lay graph"
done; done
for lay in stream_major time_major; do
timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 1 --steps 100 --warmup 5 --layout $lay 2>>gpurun_out/s1.err | show "S=1 $lay eager"
timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 4 --steps 100 --warmup 5 --graph --layout $lay 2>>gpurun_out/s1.err | show "S=1 h4 $lay graph"
done
tail -n 3 gpurun_out/s1.err
