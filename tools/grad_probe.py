"""E8-full gradient parity probe: per-tensor relative errors (product vs oracle autograd in fp64) for several clip lengths."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
import cleanumamba_oracle as orc
from cleanumamba_b200.network import Net
sums = json.load(open(os.path.join(ROOT, "tests/golden/full_init_seed0_sums.json")))["DNS-CleanUMamba-3N-E8"]
for secs in [float(a) for a in sys.argv[1:]] or [1.0]:
    torch.manual_seed(0)
    base = Net("CleanUMamba", dict(sums["config"], math_mode="fp32"))
    clean, noisy = orc.synth_batch(1, secs, seed=41)
    ref = {}
    for dt in (torch.float32, torch.float64):
        sd = {k: v.detach().clone().to(dt).requires_grad_() for k, v in base.state_dict().items()}
        out = orc.forward(sd, noisy.to(dt), differentiable=True, dtype=dt)
        (F.l1_loss(out, clean.to(dt)) + (out ** 2).mean()).backward()
        ref[dt] = {k: v.grad.double() for k, v in sd.items()}
    net = base.cuda().train()
    out = net(noisy.clone().cuda())
    (F.l1_loss(out, clean.cuda()) + (out ** 2).mean()).backward()
    rows = []
    for k, p in net.named_parameters():
        g64, g32 = ref[torch.float64][k], ref[torch.float32][k]
        s = g64.abs().max().item()
        rows.append(((p.grad.double().cpu() - g64).abs().max().item() / s, (g32 - g64).abs().max().item() / s, k))
    rows.sort(reverse=True)
    print(f"== {secs} s: product(fp32 kernels) vs oracle fp64 | oracle fp32 vs fp64")
    for r in rows[:6]:
        print("   %.3e | %.3e  %s" % r)
    k = "encoder.7.0.bias"
    d = (net.state_dict()[k] * 0)  # placeholder
    g = dict(net.named_parameters())[k].grad.double().cpu()
    diff = (g - ref[torch.float64][k]).abs()
    top = diff.topk(5)
    print("   enc7.0.bias top diffs", [(int(i), float(v), float(ref[torch.float64][k][i])) for v, i in zip(top.values, top.indices)])
