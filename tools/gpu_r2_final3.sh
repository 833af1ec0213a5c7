#!/bin/bash
# refresh of the N=1 default line and the single-stream record after the small-M GEMM path (last code change of round 2)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 1 --steps 100 --warmup 5 --graph > gpurun_out/final2_bench_stream_s1_h1_graph.json 2>> gpurun_out/final2_stream.err
timeout 300 python bench.py --mode stream --model e6 --streams 4 --hops 1 --steps 100 --warmup 5 --graph > gpurun_out/final2_bench_stream_s4_h1_graph.json 2>> gpurun_out/final2_stream.err
python - <<PY
import json
d=json.loads(open('gpurun_out/final2_bench_n1.json').read().strip().splitlines()[-1])
print('N=1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
t=d.get('train',{}); print('train', t.get('value'), t.get('ms_per_step'), t.get('error'))
for k,v in (d.get('stream') or {}).items(): print('stream', k, v.get('value'), v.get('ms_per_step'), v.get('config',{}).get('real_time_factor_per_stream')) if isinstance(v,dict) else print(k,v)
for f in ('s1','s4'):
    r=json.loads(open(f'gpurun_out/final2_bench_stream_{f}_h1_graph.json').read().strip().splitlines()[-1]); print(f, r['ms_per_step'], r['config']['real_time_factor_per_stream'], r['config']['buffer_layout'])
PY
