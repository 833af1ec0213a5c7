#!/bin/bash
# round-2 FINAL measurements (after the time-major streaming session / TMA-staged state update) on N GPUs of one box: default bench
# line (offline + train + stream sub-records) and, at N=1, the reference arm, sweep, training, pruned configuration, streaming variants
# and the ncu launch list of one streaming call
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
N=${N:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; echo "bench rc=$?"
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final2_bench_reference.json 2> gpurun_out/final2_bench_reference.err; echo "ref rc=$?"
  timeout 600 python bench.py --mode sweep > gpurun_out/final2_bench_sweep.json 2> gpurun_out/final2_bench_sweep.err; echo "sweep rc=$?"
  timeout 300 python bench.py --mode train --steps 10 > gpurun_out/final2_bench_train.json 2> gpurun_out/final2_bench_train.err; echo "train rc=$?"
  timeout 300 python bench.py --mode pruned > gpurun_out/final2_bench_pruned.json 2> gpurun_out/final2_bench_pruned.err; echo "pruned rc=$?"
  for h in 1 2 4 16 64; do
    timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops $h --steps 20 --warmup 5 > gpurun_out/final2_bench_stream_h$h.json 2>> gpurun_out/final2_stream.err
    timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops $h --steps 20 --warmup 5 --graph > gpurun_out/final2_bench_stream_h${h}_graph.json 2>> gpurun_out/final2_stream.err
  done
  for h in 1 2 4; do
    timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops $h --steps 20 --warmup 5 --graph --state-f16 > gpurun_out/final2_bench_stream_h${h}_graph_f16state.json 2>> gpurun_out/final2_stream.err
  done
  timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 20 --warmup 5 --graph --layout stream_major > gpurun_out/final2_bench_stream_h1_graph_stream_major.json 2>> gpurun_out/final2_stream.err
  timeout 300 python bench.py --mode stream --model e6 --streams 1 --hops 1 --steps 100 --warmup 5 --graph > gpurun_out/final2_bench_stream_s1_h1_graph.json 2>> gpurun_out/final2_stream.err
  timeout 300 python tools/pruned_probe.py > gpurun_out/final2_pruned.log 2>&1; tail -8 gpurun_out/final2_pruned.log
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:"_kernel" -c 3000 --csv --log-file gpurun_out/launches_stream_h1.csv \
      python -u bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 2 --warmup 3 > gpurun_out/ncu_launches_stream.log 2>&1; echo "ncu launches rc=$?"
  for f in gpurun_out/final2_bench_stream_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('final2_bench_')[1], d['ms_per_step'], d['config']['real_time_factor_per_stream'], d['config'].get('buffer_layout'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
  done
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final2_bench_n$N.json 2> gpurun_out/final2_bench_n$N.err; echo "bench rc=$?"; tail -n 2 gpurun_out/final2_bench_n$N.err
fi
python - <<PY
import json
d=json.loads(open('gpurun_out/final2_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
t=d.get('train',{}); print('train', t.get('value'), t.get('ms_per_step'), json.dumps(t.get('config',{}).get('grad_allreduce')), t.get('error'))
for k,v in (d.get('stream') or {}).items(): print('stream', k, v.get('value'), v.get('ms_per_step'), v.get('config',{}).get('real_time_factor_per_stream')) if isinstance(v,dict) else print(k,v)
PY
