"""Micro-benchmark of cum_selective_scan_fwd at the E8-full shapes (B x 624 tokens, d_inner 2048, d_state 64)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleanumamba_b200 import _lib  # noqa: E402

B, L, D, N, R = int(os.environ.get("B", 64)), 624, 2048, 64, 32
dev = torch.device("cuda:0")
lib = _lib.init(dev)
g = torch.Generator(device="cuda").manual_seed(0)
xz = torch.randn(B, L, 2 * D, device=dev, generator=g)
xc = torch.randn(B, L, D, device=dev, generator=g)
dt = torch.randn(B, L, D, device=dev, generator=g) * 0.5
xdbl = torch.randn(B, L, R + 2 * N, device=dev, generator=g)
y = torch.empty(B, L, D, device=dev)
a2 = -torch.exp(torch.randn(D, N, device=dev, generator=g) * 0.5) * 1.4427
Dk, bias = torch.randn(D, device=dev, generator=g), torch.randn(D, device=dev, generator=g) * 0.5 - 2
s = _lib.ScanDesc()
s.u, s.u_bs, s.u_rs = xc.data_ptr(), L * D, D
s.delta, s.dl_bs, s.dl_rs = dt.data_ptr(), L * D, D
s.z, s.z_bs, s.z_rs = xz.data_ptr() + 4 * D, L * 2 * D, 2 * D
ld = R + 2 * N
s.Bm, s.B_bs, s.B_rs = xdbl.data_ptr() + 4 * R, L * ld, ld
s.Cm, s.C_bs, s.C_rs = xdbl.data_ptr() + 4 * (R + N), L * ld, ld
s.y, s.y_bs, s.y_rs = y.data_ptr(), L * D, D
s.a2, s.Dskip, s.delta_bias, s.h0, s.h_out, s.h_ckpt = a2.data_ptr(), Dk.data_ptr(), bias.data_ptr(), 0, 0, 0
s.batch, s.len, s.d, s.n_state, s.delta_softplus = B, L, D, N, 1
for _ in range(3):
    _lib.check(lib.cum_selective_scan_fwd(C.byref(s), _lib.stream_ptr()), "scan")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.cum_selective_scan_fwd(C.byref(s), _lib.stream_ptr())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
upd = B * L * D * N
print(f"scan fwd B={B}: {ms:.3f} ms  {upd / ms / 1e9:.2f} G updates/ms-> {upd / (ms / 1e3) / 1e12:.2f}e12 upd/s ({upd / (ms / 1e3) / 4.5e12 * 100:.0f}% of MUFU ceiling)  "
      f"{4 * B * L * (4 * D + 2 * N) / ms / 1e6:.0f} GB/s algorithmic  checksum {float(y.double().abs().mean()):.6f}")
