#!/bin/bash
# round 2, time-major streaming session: new tests, then 1-hop / 16-hop stream benches in both layouts + the fp16-state variant
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_stream_tm.py -m gpu -rP --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_tm.log 2>&1; echo "pytest tm rc=$?"
grep -E "^\[|passed|failed|^E  |Error" gpurun_out/tests_tm.log | tail -20
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config']['real_time_factor_per_stream'], d['config'].get('buffer_layout'), d['gpu_launches'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"; }
for lay in stream_major time_major; do
  timeout 300 python bench.py --mode stream --model e6 --streams 4096 --hops 1 --steps 20 --warmup 5 --layout $lay 2>gpurun_out/stream_$lay.err | tee gpurun_out/stream_h1_$lay.json | show "h1 $lay eager"
  timeout 300 python bench.py --mode stream --model e6 --streams 4096 --hops 1 --steps 20 --warmup 5 --layout $lay --graph 2>>gpurun_out/stream_$lay.err | tee gpurun_out/stream_h1_${lay}_graph.json | show "h1 $lay graph"
done
timeout 300 python bench.py --mode stream --model e6 --streams 4096 --hops 1 --steps 20 --warmup 5 --graph --state-f16 2>gpurun_out/stream_f16.err | tee gpurun_out/stream_h1_f16state_graph.json | show "h1 f16state graph"
timeout 300 python bench.py --mode stream --model e6 --streams 4096 --hops 16 --steps 10 --warmup 3 --layout time_major 2>>gpurun_out/stream_time_major.err | tee gpurun_out/stream_h16_time_major.json | show "h16 tm"
timeout 300 python bench.py --mode stream --model e6 --streams 4096 --hops 16 --steps 10 --warmup 3 --layout stream_major 2>>gpurun_out/stream_stream_major.err | tee gpurun_out/stream_h16_stream_major.json | show "h16 sm"
tail -3 gpurun_out/stream_*.err
