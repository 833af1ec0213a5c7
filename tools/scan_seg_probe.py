import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import cleanumamba_oracle as orc
from cleanumamba_b200 import ops
for (b, d, l, n, with_h0) in [(1, 256, 2000, 64, True), (1, 256, 2000, 64, False), (1, 256, 640, 64, True)]:
    g = torch.Generator().manual_seed(b * 1000 + d + l + n)
    u = torch.randn(b, d, l, generator=g)
    delta = torch.randn(b, d, l, generator=g) * 0.5
    A = -torch.exp(torch.randn(d, n, generator=g) * 0.5 + 0.5)
    Bm, Cm = torch.randn(b, n, l, generator=g), torch.randn(b, n, l, generator=g)
    D, z = torch.randn(d, generator=g), torch.randn(b, d, l, generator=g)
    bias = torch.randn(d, generator=g) * 0.5 - 2.0
    h0 = torch.randn(b, d, n, generator=g) if with_h0 else None
    y_ref, h_ref = orc.selective_scan(u, delta, A, Bm, Cm, D, z, bias, True, h0=h0, return_last_state=True)
    cu = lambda t: None if t is None else t.cuda()
    args = (cu(u), cu(delta), cu(A), cu(Bm), cu(Cm), cu(D), cu(z), cu(bias), True)
    y, h = ops.selective_scan_fn(*args, return_last_state=True, initial_state=cu(h0))
    y1, h1 = ops.selective_scan_fn(*args, return_last_state=True, initial_state=cu(h0), segment_parallel=False)
    import ctypes as C
    from cleanumamba_b200 import _lib
    sd = _lib.ScanDesc(); sd.batch, sd.len, sd.d, sd.n_state = b, l, d, n
    print("   workspace bytes", _lib.load().cum_selective_scan_workspace_bytes(C.byref(sd)), "max|y_seg - y_seq|", (y - y1).abs().max().item(),
          "max|h_seg - h_seq|", (h - h1).abs().max().item())
    e_seg = (y.cpu() - y_ref).abs().amax(dim=(0, 1))
    e_seq = (y1.cpu() - y_ref).abs().amax(dim=(0, 1))
    print((b, d, l, n, with_h0), "seg err by t-block:", [f"{e_seg[i:i+80].max().item():.1e}" for i in range(0, l, 240)])
    print("   seq err by t-block:", [f"{e_seq[i:i+80].max().item():.1e}" for i in range(0, l, 240)], "h:", (h.cpu()-h_ref).abs().max().item(), (h1.cpu()-h_ref).abs().max().item())
