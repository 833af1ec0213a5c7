"""torch.profiler view of one training step (E8 full, 16 x 10 s): CUDA time by kernel name, so the PyTorch-side work (loss, Adam,
gradient unpacking) shows next to our kernels.  Usage: python tools/train_profile.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cleanumamba_b200.network import Net
from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG, loss_fn
from cleanumamba_b200.fused_loss import FusedMultiResolutionSTFTLoss
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Net("CleanUMamba", dict(bench.CONFIGS["e8"], math_mode="f16x3")).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=2e-4, fused=True)
mr = FusedMultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)
noisy = bench.synth_noisy(16, 10.0, 1234).to(dev)
clean = bench.synth_noisy(16, 10.0, 99).to(dev) * 0.5
work = torch.empty_like(noisy)

def step():
    work.copy_(noisy)
    opt.zero_grad(set_to_none=True)
    loss, _ = loss_fn(net, (clean, work), mrstftloss=mr)
    loss.backward()
    opt.step()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
