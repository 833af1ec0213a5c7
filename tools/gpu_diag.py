"""Staged GPU bring-up diagnostic: each step prints + flushes, so a hang is attributable from gpurun_out/diag.log."""
import faulthandler
import os
import sys
import time

faulthandler.enable()
faulthandler.dump_traceback_later(240, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
T0 = time.time()


def say(*a):
    print(f"[{time.time() - T0:7.2f}s]", *a, flush=True)


say("nproc", os.cpu_count())
import torch  # noqa: E402
say("torch imported", torch.__version__, "threads", torch.get_num_threads())
torch.set_num_threads(min(16, os.cpu_count()))
assert torch.cuda.is_available()
say("device", torch.cuda.get_device_name(0))
x = torch.randn(1024, 1024, device="cuda")
torch.cuda.synchronize()
say("torch cuda ok")
from cleanumamba_b200 import _lib, ops  # noqa: E402
lib = _lib.init(torch.device("cuda:0"))
say("cum_init ok")
import torch.nn.functional as F  # noqa: E402
import cleanumamba_oracle as orc  # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
# normalise
xs = (torch.randn(3, 1000, generator=g) * 0.3).to(dev)
std = torch.empty(3, device=dev)
_lib.check(lib.cum_wave_normalize_fwd(xs.data_ptr(), std.data_ptr(), 3, 1000, _lib.stream_ptr()), "norm")
torch.cuda.synchronize(); say("normalize ok", std.tolist())
# LN
y, r = ops.layer_norm_residual(torch.randn(50, 512, device=dev), torch.randn(50, 512, device=dev), torch.ones(512, device=dev), torch.zeros(512, device=dev))
torch.cuda.synchronize(); say("ln ok", float(y.abs().max()))
# dwconv
xx = torch.randn(2, 64, 70, generator=g); w = torch.randn(64, 4, generator=g); b = torch.randn(64, generator=g)
yy = ops.causal_conv1d_fn(xx.to(dev), w.to(dev), b.to(dev), "silu")
torch.cuda.synchronize(); say("dwconv ok", float((yy.cpu() - F.silu(F.conv1d(xx, w[:, None], b, padding=3, groups=64)[..., :70])).abs().max()))
# gemm
a = torch.randn(2, 300, 64, generator=g); wt = torch.randn(1, 128, 64, generator=g)
c = ops.gemm_bias_act(a.to(dev), wt.to(dev))
torch.cuda.synchronize(); say("gemm simt ok", float((c.cpu() - a @ wt[0].t()).abs().max()))
# scan
bb, d, l, n = 2, 64, 37, 64
u = torch.randn(bb, d, l, generator=g); dl = torch.randn(bb, d, l, generator=g) * .5
A = -torch.exp(torch.randn(d, n, generator=g) * .5); Bm = torch.randn(bb, n, l, generator=g); Cm = torch.randn(bb, n, l, generator=g)
D = torch.randn(d, generator=g); z = torch.randn(bb, d, l, generator=g); bias = torch.randn(d, generator=g) * .5 - 2
t = time.time(); yr = orc.selective_scan(u, dl, A, Bm, Cm, D, z, bias, True); say("oracle scan cpu s", time.time() - t)
cu = lambda v: v.to(dev)  # noqa: E731
ys = ops.selective_scan_fn(cu(u), cu(dl), cu(A), cu(Bm), cu(Cm), cu(D), cu(z), cu(bias), True)
torch.cuda.synchronize(); say("scan ok", float((ys.cpu() - yr).abs().max()))
# convt_out (c_pad <= 64 exercises the 16-lane groups)
for Hc in (56, 128):
    gin = torch.randn(2, Hc, 100, generator=g); wT = torch.randn(Hc, 1, 4, generator=g); sc = torch.rand(2, generator=g) + .5
    out = torch.empty(2, 1, 190, device=dev)
    _lib.check(lib.cum_convt_out_fwd(gin.permute(0, 2, 1).contiguous().to(dev).data_ptr(), 2, 100, Hc, wT[:, 0].t().contiguous().to(dev).data_ptr(),
                                     0.25, sc.to(dev).data_ptr(), 190, out.data_ptr(), 190, 0, 190, 4, 2, _lib.stream_ptr()), "convt_out")
    torch.cuda.synchronize()
    ref = (F.conv_transpose1d(gin, wT, None, stride=2) + 0.25)[..., :190] * sc[:, None, None]
    say("convt_out ok", Hc, float((out.cpu() - ref).abs().max()))
# whole model, pruned checkpoint fixture
import json  # noqa: E402
from cleanumamba_b200.network import Net  # noqa: E402
fx = torch.load(os.path.join(ROOT, "tests/golden/e8_pruned_500k.pt"), map_location="cpu", weights_only=True)
net = Net("CleanUMamba", json.loads(fx["config"])); net.load_pruned_state_dict(fx["state_dict"]); net = net.cuda().float().eval()
say("model built")
with torch.no_grad():
    yv = net(fx["noisy"].cuda())
torch.cuda.synchronize()
say("model forward ok, max-abs vs reference golden", float((yv.cpu() - fx["denoised"]).abs().max()), "rms", float(fx["denoised"].pow(2).mean().sqrt()))
if "--tc" in sys.argv:
    for math in ("tf32", "tf32x3", "bf16x3", "f16x3"):
        for (bt, rows, k, n_) in ((1, 300, 64, 128), (2, 1000, 768, 768), (3, 130, 104, 200)):
            a = torch.randn(bt, rows, k, generator=g); wt = torch.randn(1, n_, k, generator=g) / k ** .5
            c = ops.gemm_bias_act(a.to(dev), wt.to(dev), math=math)
            torch.cuda.synchronize()
            ref = (a.double() @ wt[0].double().t())
            say("gemm", math, (bt, rows, k, n_), "max err", float((c.cpu().double() - ref).abs().max()), "ref max", float(ref.abs().max()))
say("DONE")
