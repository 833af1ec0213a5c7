#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])
print('train', json.dumps(d.get('train'))[:1500])
print('stream', json.dumps(d.get('stream'))[:2500])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
