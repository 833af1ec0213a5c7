#!/bin/bash
# 2-GPU call: NCCL gradient-equality test, then the default bench line at N=2 (offline + train + stream sub-records)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-8}
N=${N:-2}
nvidia-smi -L
timeout 600 python -u -m pytest tests/test_gpu_dist.py -m gpu -rP --timeout 300 -q -p no:cacheprovider > gpurun_out/tests_dist.log 2>&1; echo "pytest dist rc=$?"; tail -5 gpurun_out/tests_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
t=d.get('train',{}); print('train', t.get('value'), t.get('ms_per_step'), json.dumps(t.get('config',{}).get('grad_allreduce')), t.get('error'))
s=d.get('stream',{})
for k,v in s.items(): print('stream', k, v.get('value'), v.get('ms_per_step'), v.get('config',{}).get('real_time_factor_per_stream')) if isinstance(v,dict) else print(k,v)
PY
