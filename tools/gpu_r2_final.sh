#!/bin/bash
# round-2 final measurements on N GPUs of one box: default bench line (offline + train + stream sub-records) and, at N=1, the
# reference arm, the clip-length sweep, the pruned-checkpoint configuration and the streaming variants
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
N=${N:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"
  timeout 600 python bench.py --mode sweep > gpurun_out/final_bench_sweep.json 2> gpurun_out/final_bench_sweep.err; echo "sweep rc=$?"
  timeout 300 python bench.py --mode train --steps 10 > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err; echo "train rc=$?"
  for h in 1 4 16 64; do timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops $h --steps 10 > gpurun_out/final_bench_stream_h$h.json 2> gpurun_out/final_bench_stream_h$h.err; done
  timeout 300 python tools/pruned_probe.py > gpurun_out/final_pruned.log 2>&1; tail -8 gpurun_out/final_pruned.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench_n$N.err
fi
python - <<PY
import json
d=json.loads(open('gpurun_out/final_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
t=d.get('train',{}); print('train', t.get('value'), t.get('ms_per_step'), json.dumps(t.get('config',{}).get('grad_allreduce')), t.get('error'))
for k,v in (d.get('stream') or {}).items(): print('stream', k, v.get('value'), v.get('ms_per_step'), v.get('config',{}).get('real_time_factor_per_stream')) if isinstance(v,dict) else print(k,v)
PY
