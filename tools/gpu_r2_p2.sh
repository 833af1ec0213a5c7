#!/bin/bash
# time-major session for irregular channel counts (pruned checkpoints): tests, then many-stream timings of the pruned models in both layouts
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_stream_tm.py tests/test_gpu_stream.py -m gpu --timeout 300 -x -q -p no:cacheprovider > gpurun_out/tests_p2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/tests_p2.log | tail -12
python - <<'P'
import json, time, torch, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
from conftest import load_golden
from cleanumamba_b200.network import Net
for name in ("e6_pruned_200k", "e8_pruned_500k"):
    fx = load_golden(name)
    net = Net("CleanUMamba", {**json.loads(fx["config"]), "math_mode": "f16x3"})
    net.load_pruned_state_dict(fx["state_dict"]); net = net.cuda().float().eval()
    hop, fl = net.total_stride, net.frame_length
    for S in (512, 4096):
        for layout in ("stream_major", "time_major"):
            for graph in (False, True):
                sess = net.stream_session(batch=S, layout=layout)
                sess.feed(torch.randn(S, fl - hop, device="cuda") * 0.1)
                chunk = torch.randn(S, hop, device="cuda") * 0.1
                for _ in range(5): sess.feed(chunk)
                if graph: sess.capture_graph(hop); sess.feed(chunk)
                torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(30): sess.feed(chunk)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 30
                print(f"{name} streams={S} {layout} graph={graph}: {ms:.3f} ms per 1-hop call ({hop / 16.0 / ms:.2f}x real time)", flush=True)
                del sess
P
timeout 300 python bench.py --mode stream --model e8 --streams 1024 --hops 1 --steps 20 --warmup 5 --graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('E8 full 1024 streams h1 graph', d['ms_per_step'], d['config']['real_time_factor_per_stream'], d['config']['buffer_layout'])"
