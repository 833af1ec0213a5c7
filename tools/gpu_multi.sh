#!/bin/bash
# N-GPU runs (utterance-sharded offline forward; data-parallel training step with bucketed NCCL all-reduce)
N=${N:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/bench_offline_n$N.json 2> gpurun_out/bench_offline_n$N.err; echo "offline n=$N rc=$?"; tail -c 1200 gpurun_out/bench_offline_n$N.json | head -c 700; echo; tail -2 gpurun_out/bench_offline_n$N.err
timeout 600 $TR bench.py --gpus $N --mode train --steps 3 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; echo "train n=$N rc=$?"; head -c 900 gpurun_out/bench_train_n$N.json; echo; tail -2 gpurun_out/bench_train_n$N.err
timeout 600 $TR bench.py --gpus $N --mode stream --model e6 --streams 4096 --hops 16 --steps 5 --warmup 3 > gpurun_out/bench_stream_n$N.json 2> gpurun_out/bench_stream_n$N.err; echo "stream n=$N rc=$?"; head -c 600 gpurun_out/bench_stream_n$N.json; echo
timeout 300 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>/dev/null; echo "ref arm n=$N rc=$?"; head -c 300 gpurun_out/bench_ref_n$N.json
