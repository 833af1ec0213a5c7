import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cleanumamba_b200.network import Net
fx = torch.load(os.path.join(ROOT, "tests/golden/tiny_equalwidth_seed0.pt"), map_location="cpu", weights_only=True)
net = Net("CleanUMamba", {**json.loads(fx["config"]), "math_mode": "bf16"})
net.load_state_dict(fx["state_dict"]); net = net.cuda().eval()
with torch.no_grad():
    y = net(fx["noisy"].clone().cuda())
torch.cuda.synchronize()
print("bf16 tiny forward ok", float((y.cpu() - fx["denoised"]).abs().max()), float(fx["denoised"].abs().max()))
