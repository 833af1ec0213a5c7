#!/bin/bash
# End-of-round measurement suite (1 GPU): tests, headline bench (+variants, CPU baseline), reference arm, streaming hop sweep,
# training step, clip-length sweep.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -4 gpurun_out/tests.log
timeout 300 python -u -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -u bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"; tail -c 600 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 900 python -u bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; tail -c 700 gpurun_out/bench_reference.json
for H in 1 4 16 64; do
  timeout 600 python -u bench.py --mode stream --model e6 --streams 4096 --hops $H --steps 5 --warmup 3 > gpurun_out/bench_stream_h$H.json 2> gpurun_out/bench_stream_h$H.err; echo "stream hops=$H rc=$?"; tail -c 1200 gpurun_out/bench_stream_h$H.json | cut -c1-1200
done
for S in 1 256; do
  timeout 600 python -u bench.py --mode stream --model e6 --streams $S --hops 1 --steps 20 --warmup 5 > gpurun_out/bench_stream_s${S}_h1.json 2> gpurun_out/bench_stream_s${S}_h1.err; echo "stream streams=$S hops=1 rc=$?"; cut -c1-700 gpurun_out/bench_stream_s${S}_h1.json
done
timeout 600 python -u bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"; cut -c1-900 gpurun_out/bench_train.json
timeout 900 python -u bench.py --mode sweep > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo "sweep rc=$?"; grep "# sweep" gpurun_out/bench_sweep.err
