#!/bin/bash
# stream / train bench modes (1 GPU).  Output: gpurun_out/bench_<mode>_<math>.json
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -u bench.py --mode stream --model e6 --streams ${STREAMS:-4096} --hops ${HOPS:-16} --steps 5 --warmup 3 --math ${MATH:-tf32x3} > gpurun_out/bench_stream_${MATH:-tf32x3}.json 2> gpurun_out/bench_stream.err; echo "stream rc=$?"; tail -c 2500 gpurun_out/bench_stream_${MATH:-tf32x3}.json; tail -3 gpurun_out/bench_stream.err
for M in ${TRAIN_MATHS:-tf32x3 fp32}; do
timeout 600 python -u bench.py --mode train --steps 3 --warmup 3 --math $M > gpurun_out/bench_train_$M.json 2> gpurun_out/bench_train_$M.err; echo "train $M rc=$?"; tail -c 3000 gpurun_out/bench_train_$M.json; tail -3 gpurun_out/bench_train_$M.err
done
