import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cleanumamba_b200.network import Net
torch.manual_seed(0)
net = Net("CleanUMamba", dict(bench.CONFIGS["e8"])).cuda().eval()
eng = net.engine()
with torch.no_grad():
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
        print(f"E8-full 1 x {sec:g} s: {e0.elapsed_time(e1)/5:.3f} ms/forward; scan {prof['selective_scan']['ms']/5:.3f} ms ({prof['selective_scan']['launches']//5} scan calls per forward)")
