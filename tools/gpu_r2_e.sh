#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 600 python -u -m pytest tests/test_gpu_ops.py -m gpu -rP --timeout 120 -x -q -p no:cacheprovider -k "scan" > gpurun_out/tests_scan.log 2>&1; echo "pytest scan rc=$?"; tail -5 gpurun_out/tests_scan.log
timeout 300 python tools/bench_scan_small.py 2>&1 | tee gpurun_out/bench_scan_small.log
# streaming at 4096 streams x 1 hop: eager vs CUDA-graph replay
for g in "" "--graph"; do timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 20 $g > gpurun_out/bench_stream_h1$g.json 2> gpurun_out/bench_stream_h1$g.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_stream_h1$g.json').read().strip().splitlines()[-1])
print('stream h1 $g', d['value'], d['ms_per_step'], d['config']['real_time_factor_per_stream'], d['gpu_launches'])
PY
done
# long-clip forward at batch 1 (segment-parallel scan inside the model)
timeout 300 python - <<PY
import torch, sys, json
sys.path.insert(0,'.')
import bench
from cleanumamba_b200.network import Net
torch.manual_seed(0)
net = Net("CleanUMamba", dict(bench.CONFIGS["e8"])).cuda().eval()
eng = net.engine()
for sec in (10.0, 60.0):
    x = bench.synth_noisy(1, sec, 5).cuda()
    w = torch.empty_like(x)
    for _ in range(3):
        w.copy_(x); net(w)
    torch.cuda.synchronize()
    eng.prof = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        w.copy_(x); net(w)
    e1.record(); torch.cuda.synchronize()
    prof = eng.profile_summary(); eng.prof = None
    print(f"E8-full 1 x {sec:g} s: {e0.elapsed_time(e1)/5:.3f} ms/forward; scan {prof['selective_scan']['ms']/5:.3f} ms ({prof['selective_scan']['launches']//5} launches)")
PY
