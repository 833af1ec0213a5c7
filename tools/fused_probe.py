"""Which fused end kernel breaks with many tiles per CTA?  Compares each against PyTorch on a 4 x 10 s sized problem."""
import ctypes as C, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
import cleanumamba_oracle as orc
from cleanumamba_b200 import _lib
dev = "cuda"
lib = _lib.init(torch.device(dev))
def split(w):
    a = w.abs().max().item(); e = int(math.floor(math.log2(8.0 / a)))
    hi, lo = torch.empty_like(w, dtype=torch.float16), torch.empty_like(w, dtype=torch.float16)
    _lib.check(lib.cum_split_f16(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), float(2.0 ** e), _lib.stream_ptr()), "split")
    return hi, lo, float(2.0 ** -e)
g = torch.Generator().manual_seed(0)
b, length = 4, 160000
x = torch.randn(b, 1, length, generator=g)
w0, b0 = torch.randn(64, 1, 4, generator=g) * 0.5, torch.randn(64, generator=g) * 0.1
w1, b1 = torch.randn(128, 64, generator=g) / 8, torch.randn(128, generator=g) * 0.1
rows = (length - 4) // 2 + 1
ref = orc.glu(F.conv1d(F.relu(F.conv1d(x, w0, b0, stride=2)), w1[:, :, None], b1)).permute(0, 2, 1)
wi = torch.stack([w1[:64], w1[64:]], 1).reshape(128, 64).contiguous().to(dev)
bi = torch.stack([b1[:64], b1[64:]], 1).reshape(128).contiguous().to(dev)
hi, lo, inv = split(wi)
xd = x[:, 0].contiguous().to(dev); cw, cb = w0[:, 0, :].t().contiguous().to(dev), b0.to(dev)
for rep in range(3):
    out = torch.full((b, rows, 64), float("nan"), device=dev)
    d = _lib.Enc0BlockDesc()
    d.x, d.x_stride, d.batch, d.length = xd.data_ptr(), length, b, length
    d.conv_w, d.conv_b, d.glu_w_hi, d.glu_w_lo, d.glu_b = cw.data_ptr(), cb.data_ptr(), hi.data_ptr(), lo.data_ptr(), bi.data_ptr()
    d.acc_scale, d.w_lo_is_zero, d.out, d.rows_out, d.channels, d.channels_out = inv, 0, out.data_ptr(), rows, 64, 64
    _lib.check(lib.cum_enc0_block_fwd(C.byref(d), _lib.stream_ptr()), "enc0")
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().amax(dim=2)        # (b, rows)
    bad = (err > 1e-4)
    tiles = bad.view(b, -1)[:, : (rows // 128) * 128].view(b, -1, 128).any(2)
    idx = tiles.nonzero()
    print(f"enc0 rep {rep}: max err {err.max().item():.3e}; bad rows {int(bad.sum())}; bad tiles {len(idx)} of {tiles.numel()}; first bad (clip, tile): {idx[:6].tolist()}",
          "nan" if torch.isnan(out).any() else "")
    if len(idx):
        c, t = idx[0].tolist()
        e = err[c, t * 128:(t + 1) * 128]
        print("   rows of first bad tile with err>1e-4:", (e > 1e-4).nonzero().flatten()[:20].tolist(), "tile linear index", c * ((rows + 127) // 128) + t, "-> CTA", (c * ((rows + 127) // 128) + t) % 148, "iteration", (c * ((rows + 127) // 128) + t) // 148)
# dec_last
a = torch.randn(b, 64, rows, generator=g)
wt, bt = torch.randn(64, 1, 4, generator=g) / 8, torch.randn(1, generator=g) * 0.1
full = F.conv_transpose1d(orc.glu(F.conv1d(a, w1[:, :, None], b1)), wt, bt, stride=2)
L = 2 * rows + 2
acl = a.permute(0, 2, 1).contiguous().to(dev); tw = wt[:, 0, :].t().contiguous().to(dev)
for rep in range(3):
    out = torch.full((b, 1, L), float("nan"), device=dev)
    d = _lib.DecLastBlockDesc()
    d.a, d.batch, d.rows_in = acl.data_ptr(), b, rows
    d.glu_w_hi, d.glu_w_lo, d.glu_b, d.acc_scale, d.w_lo_is_zero = hi.data_ptr(), lo.data_ptr(), bi.data_ptr(), inv, 0
    d.convt_w, d.convt_bias, d.scale = tw.data_ptr(), float(bt[0]), 0
    d.out, d.out_stride, d.out_length, d.channels, d.channels_gated = out.data_ptr(), L, L, 64, 64
    _lib.check(lib.cum_dec_last_block_fwd(C.byref(d), _lib.stream_ptr()), "dec_last")
    torch.cuda.synchronize()
    err = (out.cpu() - full).abs()[:, 0]
    bad = err > 1e-4
    print(f"dec_last rep {rep}: max err {err.max().item():.3e}; bad samples {int(bad.sum())}; first bad: {bad.nonzero()[:5].tolist()}", "nan" if torch.isnan(out).any() else "")
