"""Per-call timing of one training step (E8 full, 16 x 10 s): every C-ABI call in launch order with its CUDA-event time and,
for the GEMMs, the algorithmic TFLOP/s.  Usage: python tools/train_calls.py [kind-filter]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cleanumamba_b200.network import Net
from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG, loss_fn
from cleanumamba_b200.fused_loss import FusedMultiResolutionSTFTLoss

flt = sys.argv[1] if len(sys.argv) > 1 else ""
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Net("CleanUMamba", dict(bench.CONFIGS["e8"], math_mode="f16x3")).to(dev).train()
mr = FusedMultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)
noisy = bench.synth_noisy(16, 10.0, 1234).to(dev)
clean = bench.synth_noisy(16, 10.0, 99).to(dev) * 0.5
eng = net.train_engine()
for it in range(3):
    eng.prof = [] if it == 2 else None
    net.zero_grad(set_to_none=True)
    loss, _ = loss_fn(net, (clean, noisy.clone()), mrstftloss=mr)
    loss.backward()
torch.cuda.synchronize()
tot = {}
for i, (kind, e0, e1, fl, nb) in enumerate(eng.prof):
    ms = e0.elapsed_time(e1)
    tot[kind] = tot.get(kind, 0.0) + ms
    if flt in kind:
        print(f"{i:4d} {kind:20s} {ms*1e3:9.1f} us  {fl/1e9:9.2f} GFLOP  {fl/ms/1e9 if ms else 0:8.1f} TFLOP/s")
print({k: round(v, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}, "sum", round(sum(tot.values()), 3))
