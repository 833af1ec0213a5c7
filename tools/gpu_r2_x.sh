#!/bin/bash
# full GPU suite + smoke + default bench line (regression check after the ABI v3 / gemm producer changes)
bash tools/gpu_tests.sh
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1])
print('offline', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print('train', d['train'].get('ms_per_step'), d['train'].get('error'))
for k,v in d['stream'].items():
    print(k, v.get('ms_per_step') if isinstance(v,dict) else v, v.get('config',{}).get('real_time_factor_per_stream') if isinstance(v,dict) else '')
P
tail -n 3 gpurun_out/bench_x.err
