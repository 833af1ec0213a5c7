"""Per-kernel CUDA time of the E8-pruned-500K forward at batch 1 x 10 s (eager and graph replay), torch.profiler."""
import json, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from conftest import load_golden  # noqa: E402
from cleanumamba_b200.network import Net  # noqa: E402
from torch.profiler import profile, ProfilerActivity

fx = load_golden("e8_pruned_500k")
net = Net("CleanUMamba", json.loads(fx["config"]))
net.load_pruned_state_dict(fx["state_dict"])
net = net.cuda().float().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.randn(B, 1, 160000, device="cuda") * 0.1
w = torch.empty_like(x)
eng = net.engine() if hasattr(net, "engine") else None
with torch.no_grad():
    for _ in range(5):
        w.copy_(x); net(w)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            w.copy_(x); net(w)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
evs = [e for e in prof.events() if e.device_type.name == "CUDA"]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 10
print("kernels per forward:", n)
seg = evs[-n:]
t0 = seg[0].time_range.start
for e in seg:
    print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:6.1f} us  {e.name[:80]}")
