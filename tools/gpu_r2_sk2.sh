#!/bin/bash
# small-M path: threshold A/B (columns x streams per level) at a few streams; single-stream latency of the shipped pruned checkpoints
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['ms_per_step'], d['config'].get('real_time_factor_per_stream'))"; }
for S in 4 16 64; do for T in 64 256 1024; do
CUM_STREAM_SMALL_ROWS=$T timeout 300 python bench.py --mode stream --model e6 --streams $S --hops 1 --steps 100 --warmup 5 --graph 2>>gpurun_out/sk2.err | show "S=$S small_rows=$T"
done; done
python - <<'P'
import json, torch, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
from conftest import load_golden
from cleanumamba_b200.network import Net
import os
for name in ("e6_pruned_200k", "e8_pruned_500k", "mini_mamba_442k"):
    fx = load_golden(name)
    for sk in ("1", "0"):
        net = Net("CleanUMamba", {**json.loads(fx["config"]), "math_mode": "f16x3"})
        net.load_pruned_state_dict(fx["state_dict"]); net = net.cuda().float().eval()
        hop, fl = net.total_stride, net.frame_length
        sess = net.stream_session(batch=1)
        if sk == "0":
            sess.SMALL_ROWS = -1
        sess.feed(torch.randn(1, fl - hop, device="cuda") * 0.1)
        chunk = torch.randn(1, hop, device="cuda") * 0.1
        for _ in range(5): sess.feed(chunk)
        sess.capture_graph(hop); sess.feed(chunk)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): sess.feed(chunk)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        print(f"{name} single stream, 1 hop ({hop} samples) per call, graph, small-M path={'on' if sk == '1' else 'off'}: {ms:.3f} ms ({hop / 16.0 / ms:.1f}x real time)", flush=True)
P
tail -n 3 gpurun_out/sk2.err
