"""Timing of the offline forward on the shipped pruned checkpoint geometry (BASELINE configs[0]: E8-pruned-500K, 4 x 10 s)."""
import json
import sys
import time

import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from conftest import load_golden  # noqa: E402
from cleanumamba_b200.network import Net  # noqa: E402

for name in ("e8_pruned_500k", "e6_pruned_200k", "mini_mamba_442k"):
    fx = load_golden(name)
    net = Net("CleanUMamba", json.loads(fx["config"]))
    net.load_pruned_state_dict(fx["state_dict"])
    net = net.cuda().float().eval()
    for B in (1, 4, 64):
        x = torch.randn(B, 1, 160000, device="cuda") * 0.1
        w = torch.empty_like(x)
        with torch.no_grad():
            for _ in range(3):
                w.copy_(x); net(w)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                w.copy_(x); net(w)
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 10 * 1e3
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} B={B}: {ms:.3f} ms/forward (wall {wall:.3f}) -> {B * 10 / (ms / 1e3):.0f} audio-s/s", flush=True)
