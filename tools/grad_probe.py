"""E8-full gradient parity probe: per-tensor relative errors (product vs oracle autograd) in two math modes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
import cleanumamba_oracle as orc
from cleanumamba_b200.network import Net
sums = json.load(open(os.path.join(ROOT, "tests/golden/full_init_seed0_sums.json")))["DNS-CleanUMamba-3N-E8"]
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
res = {}
for math in ("fp32", "f16x3", "tf32x3"):
    torch.manual_seed(0)
    net = Net("CleanUMamba", dict(sums["config"], math_mode=math))
    sd = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    clean, noisy = orc.synth_batch(1, secs, seed=41)
    if "ref" not in res:
        out_ref = orc.forward(sd, noisy, differentiable=True)
        (F.l1_loss(out_ref, clean) + (out_ref ** 2).mean()).backward()
        res["ref"] = {k: v.grad.clone() for k, v in sd.items()}
    out = net(noisy.clone().cuda())
    (F.l1_loss(out, clean.cuda()) + (out ** 2).mean()).backward()
    errs = []
    for k, p in net.named_parameters():
        gr = res["ref"][k]
        scale = gr.abs().max().item()
        d = (p.grad.cpu() - gr).abs()
        errs.append((d.max().item() / max(scale, 1e-30), k, scale, int((d > 1e-3 * scale).sum()), d.numel()))
    errs.sort(reverse=True)
    print(f"== {math}: out err {(out.detach().cpu()-out_ref.detach()).abs().max().item():.3e}")
    for e in errs[:8]:
        print("   %.3e  %-45s scale %.3e  n(>1e-3)=%d/%d" % e)
