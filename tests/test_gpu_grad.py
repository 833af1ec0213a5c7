"""GPU parity of the backward kernels: per operator against PyTorch autograd (fp32 reference of the same op), and the
whole model against autograd through the CPU oracle (== autograd through the reference's PyTorch path)."""
import ctypes as C
import json

import pytest
import torch
import torch.nn.functional as F

import cleanumamba_oracle as orc
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return ((a.detach().double().cpu() - b.detach().double().cpu()).abs().max() / b.detach().double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
def test_wgrad_taps(math):
    """fp32 CUDA-core wgrad and the tensor-core split-K wgrad (transpose + tcgen05, atomic accumulate)."""
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(DEV))
    g = torch.Generator().manual_seed(0)
    for (B, m, n, k, taps, shifts, a_rows) in [(2, 300, 128, 64, 1, (0, 0), 300), (3, 77, 72, 112, 2, (0, 1), 78),
                                                (2, 130, 200, 40, 2, (0, -1), 129), (1, 1000, 1536, 768, 1, (0, 0), 1000),
                                                (4, 5006, 768, 1024, 2, (0, 1), 5007), (2, 20000, 128, 64, 1, (0, 0), 20000)]:
        dz = torch.randn(B, m, n, generator=g)
        a = torch.randn(B, a_rows, k, generator=g)
        want = torch.zeros(taps, n, k)
        for s in range(taps):
            ash = torch.zeros(B, m, k)
            lo, hi = max(0, -shifts[s]), min(m, a_rows - shifts[s])
            ash[:, lo:hi] = a[:, lo + shifts[s]: hi + shifts[s]]
            want[s] = torch.einsum("bmn,bmk->nk", dz.double(), ash.double()).float()
        dzd, ad, dw = dz.to(DEV), a.to(DEV), torch.zeros(taps, n, k, device=DEV)
        d = _lib.WgradDesc()
        d.dz, d.dz_batch_stride, d.dz_row_stride = dzd.data_ptr(), m * n, n
        d.a, d.a_batch_stride, d.a_row_stride, d.a_rows = ad.data_ptr(), a_rows * k, k, a_rows
        d.dw, d.ldw, d.m, d.n, d.k, d.taps, d.batch = dw.data_ptr(), k, m, n, k, taps, B
        d.tap_shift[0], d.tap_shift[1] = shifts
        d.math = _lib.MATH_BY_NAME[math]
        ws = torch.empty(max(1, lib.cum_gemm_wgrad_workspace_bytes(C.byref(d)) // 4 + 1), device=DEV)
        d.workspace = ws.data_ptr()
        _lib.check(lib.cum_gemm_wgrad(C.byref(d), _lib.stream_ptr()), "wgrad")
        assert rel(dw, want) < (2e-5 if math == "fp32" else 1e-4), (B, m, n, k, taps, rel(dw, want))


def test_elementwise_backward_and_colsum():
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(DEV))
    g = torch.Generator().manual_seed(1)
    rows, H = 333, 56
    z = torch.randn(rows, 2 * H, generator=g, requires_grad=True)
    add = torch.randn(rows, H, generator=g)
    out_ref = z[:, 0::2] * torch.sigmoid(z[:, 1::2]) + add
    dout = torch.randn(rows, H, generator=g)
    out_ref.backward(dout)
    zd, addd, doutd = z.detach().to(DEV), add.to(DEV), dout.to(DEV)
    out = torch.empty(rows, H, device=DEV)
    _lib.check(lib.cum_glu_fwd(zd.data_ptr(), addd.data_ptr(), out.data_ptr(), rows, H, _lib.stream_ptr()), "glu_fwd")
    assert rel(out, out_ref) < 1e-6
    dz, db = torch.empty(rows, 2 * H, device=DEV), torch.zeros(2 * H, device=DEV)
    sc4 = torch.full((4,), -1.0, device=DEV)
    _lib.check(lib.cum_glu_bwd(zd.data_ptr(), doutd.data_ptr(), dz.data_ptr(), db.data_ptr(), rows, H, sc4.data_ptr(), _lib.stream_ptr()), "glu_bwd")
    assert rel(dz, z.grad) < 1e-5 and rel(db, z.grad.sum(0)) < 1e-5
    # the fused gradient scale: the power of two lifting max|dz| into [2^14, 2^15), and its reciprocal; == cum_grad_scale_fwd
    s, inv = sc4[0].item(), sc4[1].item()
    amax = z.grad.abs().max().item()
    assert s * inv == 1.0 and 2.0 ** 14 <= s * amax * (1 + 1e-6) and s * amax < 2.0 ** 15 * (1 + 1e-6)
    sc5 = torch.zeros(4, device=DEV)
    _lib.check(lib.cum_grad_scale_fwd(dz.data_ptr(), 0, 2 * H, 1, rows, 2 * H, sc5.data_ptr(), _lib.stream_ptr()), "grad_scale")
    assert sc5[0].item() == s and sc5[1].item() == inv
    # relu_bwd + colsum, wide rows (> 256 float4 groups)
    y = torch.relu(torch.randn(70, 1536, generator=g))
    dy = torch.randn(70, 1536, generator=g)
    yd, dyd = y.to(DEV), dy.to(DEV)
    dzr, dbr, cs = torch.empty(70, 1536, device=DEV), torch.zeros(1536, device=DEV), torch.zeros(1536, device=DEV)
    _lib.check(lib.cum_relu_bwd(yd.data_ptr(), dyd.data_ptr(), dzr.data_ptr(), dbr.data_ptr(), 70, 1536, 0, _lib.stream_ptr()), "relu_bwd")
    _lib.check(lib.cum_colsum(dyd.data_ptr(), cs.data_ptr(), 70, 1536, 0, _lib.stream_ptr()), "colsum")
    want = dy * (y > 0)
    assert torch.equal(dzr.cpu(), want) and rel(dbr, want.sum(0)) < 1e-5 and rel(cs, dy.sum(0)) < 1e-5


@pytest.mark.parametrize("rows,c", [(100, 512), (37, 114), (9, 1000)])
def test_layer_norm_backward(rows, c):
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(DEV))
    g = torch.Generator().manual_seed(rows)
    cp = (c + 7) // 8 * 8
    x = torch.randn(rows, c, generator=g, requires_grad=True)
    gamma = torch.randn(c, generator=g, requires_grad=True)
    beta = torch.randn(c, generator=g, requires_grad=True)
    dy, dres = torch.randn(rows, c, generator=g), torch.randn(rows, c, generator=g)
    (F.layer_norm(x, (c,), gamma, beta, 1e-5) * dy).sum().backward()
    pad = lambda t: F.pad(t.detach(), (0, cp - c)).contiguous().to(DEV)  # noqa: E731
    xd, dyd, drd, gd = pad(x), pad(dy), pad(dres), pad(gamma)
    dx, dg, db = torch.empty(rows, cp, device=DEV), torch.zeros(cp, device=DEV), torch.zeros(cp, device=DEV)
    _lib.check(lib.cum_ln_residual_bwd(xd.data_ptr(), dyd.data_ptr(), drd.data_ptr(), gd.data_ptr(), dx.data_ptr(), dg.data_ptr(),
                                       db.data_ptr(), 1e-5, rows, c, cp, _lib.stream_ptr()), "ln_bwd")
    assert rel(dx[:, :c], x.grad + dres) < 2e-5
    assert rel(dg[:c], gamma.grad) < 2e-5 and rel(db[:c], beta.grad) < 2e-5


@pytest.mark.parametrize("b,d,l", [(2, 64, 70), (1, 8, 3), (3, 136, 45)])
def test_dwconv_silu_backward(b, d, l):
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(DEV))
    g = torch.Generator().manual_seed(d + l)
    x = torch.randn(b, d, l, generator=g, requires_grad=True)
    w = torch.randn(d, 4, generator=g, requires_grad=True)
    bias = torch.randn(d, generator=g, requires_grad=True)
    dy = torch.randn(b, d, l, generator=g)
    (F.silu(F.conv1d(x, w[:, None], bias, padding=3, groups=d)[..., :l]) * dy).sum().backward()
    xd = x.detach().permute(0, 2, 1).contiguous().to(DEV)
    dyd = dy.permute(0, 2, 1).contiguous().to(DEV)
    wd, bd = w.detach().t().contiguous().to(DEV), bias.detach().to(DEV)
    dx, dw, db = torch.empty(b, l, d, device=DEV), torch.zeros(4, d, device=DEV), torch.zeros(d, device=DEV)
    _lib.check(lib.cum_dwconv_silu_bwd(xd.data_ptr(), l * d, d, wd.data_ptr(), bd.data_ptr(), dyd.data_ptr(), dx.data_ptr(), l * d, d,
                                       dw.data_ptr(), db.data_ptr(), b, l, d, 4, _lib.stream_ptr()), "dwconv_bwd")
    assert rel(dx.permute(0, 2, 1), x.grad) < 2e-5
    assert rel(dw.t(), w.grad) < 2e-5 and rel(db, bias.grad) < 2e-5


@pytest.mark.parametrize("b,d,l,n", [(2, 64, 37, 64), (1, 40, 50, 8), (2, 32, 16, 16), (1, 96, 33, 24)])
def test_selective_scan_backward(b, d, l, n):
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(DEV))
    g = torch.Generator().manual_seed(b + d + l + n)
    leaf = lambda *s, scale=1.0, shift=0.0: (torch.randn(*s, generator=g) * scale + shift).requires_grad_()  # noqa: E731
    u, delta, Bm, Cm, z = leaf(b, d, l), leaf(b, d, l, scale=0.5), leaf(b, n, l), leaf(b, n, l), leaf(b, d, l)
    A_log, D, bias = leaf(d, n, scale=0.5), leaf(d), leaf(d, scale=0.5, shift=-1.5)
    dout = torch.randn(b, d, l, generator=g)
    y_ref = orc.selective_scan(u, delta, -torch.exp(A_log), Bm, Cm, D, z, bias, True)
    (y_ref * dout).sum().backward()
    cl = lambda t: t.detach().permute(0, 2, 1).contiguous().to(DEV)  # noqa: E731
    ucl, dcl, zcl, Bcl, Ccl, dycl = cl(u), cl(delta), cl(z), cl(Bm), cl(Cm), cl(dout)
    a2 = (-torch.exp(A_log.detach()) * 1.4426950408889634).contiguous().to(DEV)
    Dd, biasd = D.detach().to(DEV), bias.detach().to(DEV)
    nchunks = (l + 15) // 16
    y, ck = torch.empty(b, l, d, device=DEV), torch.empty(b, nchunks, n, d, device=DEV)
    s = _lib.ScanDesc()
    s.u, s.u_bs, s.u_rs = ucl.data_ptr(), l * d, d
    s.delta, s.dl_bs, s.dl_rs = dcl.data_ptr(), l * d, d
    s.z, s.z_bs, s.z_rs = zcl.data_ptr(), l * d, d
    s.Bm, s.B_bs, s.B_rs = Bcl.data_ptr(), l * n, n
    s.Cm, s.C_bs, s.C_rs = Ccl.data_ptr(), l * n, n
    s.y, s.y_bs, s.y_rs = y.data_ptr(), l * d, d
    s.a2, s.Dskip, s.delta_bias, s.h0, s.h_out, s.h_ckpt = a2.data_ptr(), Dd.data_ptr(), biasd.data_ptr(), 0, 0, ck.data_ptr()
    s.batch, s.len, s.d, s.n_state, s.delta_softplus = b, l, d, n, 1
    _lib.check(lib.cum_selective_scan_fwd(C.byref(s), _lib.stream_ptr()), "scan_fwd")
    assert rel(y.permute(0, 2, 1), y_ref) < 2e-5
    sb = _lib.ScanBwdDesc()
    sb.fwd = s
    du, ddl, dz = (torch.empty(b, l, d, device=DEV) for _ in range(3))
    dB, dC = torch.zeros(b, l, n, device=DEV), torch.zeros(b, l, n, device=DEV)
    dA, dD, dbias = torch.zeros(d, n, device=DEV), torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    sb.h_ckpt = ck.data_ptr()
    sb.dout, sb.dout_bs, sb.dout_rs = dycl.data_ptr(), l * d, d
    sb.du, sb.du_bs, sb.du_rs = du.data_ptr(), l * d, d
    sb.ddelta, sb.ddl_bs, sb.ddl_rs = ddl.data_ptr(), l * d, d
    sb.dz, sb.dz_bs, sb.dz_rs = dz.data_ptr(), l * d, d
    sb.dB, sb.dB_bs, sb.dB_rs = dB.data_ptr(), l * n, n
    sb.dC, sb.dC_bs, sb.dC_rs = dC.data_ptr(), l * n, n
    sb.dA_log, sb.dD, sb.ddelta_bias = dA.data_ptr(), dD.data_ptr(), dbias.data_ptr()
    _lib.check(lib.cum_selective_scan_bwd(C.byref(sb), _lib.stream_ptr()), "scan_bwd")
    tol = 1e-4
    assert rel(du.permute(0, 2, 1), u.grad) < tol
    assert rel(ddl.permute(0, 2, 1), delta.grad) < tol
    assert rel(dz.permute(0, 2, 1), z.grad) < tol
    assert rel(dB.permute(0, 2, 1), Bm.grad) < tol and rel(dC.permute(0, 2, 1), Cm.grad) < tol
    assert rel(dA, A_log.grad) < tol and rel(dD, D.grad) < tol and rel(dbias, bias.grad) < tol


@pytest.mark.parametrize("name,math,seconds", [("tiny_equalwidth_seed0", "fp32", 0.2), ("e6_pruned_200k", "fp32", 0.25),
                                               ("e8_pruned_500k", "fp32", 0.3), ("e8_pruned_500k", "tf32x3", 0.3),
                                               ("e8_pruned_500k", "bf16x3", 0.3),
                                               # default mode: forward GEMMs f16x3, data / weight gradients tf32x3
                                               ("e8_pruned_500k", "f16x3", 0.3), ("mini_mamba_442k", "f16x3", 0.3)])
def test_model_gradients_match_oracle_autograd(name, math, seconds):
    from cleanumamba_b200.network import Net
    fx = load_golden(name)
    net = Net("CleanUMamba", {**json.loads(fx["config"]), "math_mode": math})
    net.load_pruned_state_dict(fx["state_dict"])
    net = net.cuda().float().train()
    T = int(16000 * seconds)
    noisy = fx["noisy"][:2, :, :T].contiguous()
    gen = torch.Generator().manual_seed(5)
    target = torch.randn(noisy.shape, generator=gen) * 0.05
    # oracle: autograd through the restated reference forward
    sd = {k: v.float().clone().requires_grad_() for k, v in fx["state_dict"].items()}
    out_ref = orc.forward(sd, noisy, differentiable=True)
    loss_ref = F.l1_loss(out_ref, target) + (out_ref ** 2).mean()
    loss_ref.backward()
    # product
    out = net(noisy.clone().cuda())
    assert out.requires_grad
    loss = F.l1_loss(out, target.cuda()) + (out ** 2).mean()
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    worst = 0.0
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        gr = sd[k].grad
        assert p.grad.shape == gr.shape, k
        scale = gr.abs().max().item()
        err = (p.grad.cpu() - gr).abs().max().item()
        if scale > 1e-12:
            worst = max(worst, err / scale)
            assert err / scale < (2e-3 if math == "fp32" else 5e-3), f"{k}: rel {err / scale:.3e}"
    print(f"\n[{name} {math}] worst per-tensor relative gradient error {worst:.3e}")


@pytest.mark.parametrize("math", ["f16x3", "fp32"])
def test_e8_full_gradients_match_oracle_autograd(math):
    """gradient parity of the FULL E8 model (41.4 M parameters, seeded random init == reference
    constructor), B = 1 x 1 s, against autograd through the CPU oracle evaluated in fp64.

    Two fp32 evaluations of this network do not agree element-wise: a ReLU pre-activation within rounding distance of 0 flips
    its mask and changes ONE output channel's gradient by O(1 %) (measured: the fp32 oracle itself is 2e-3 away from the fp64
    oracle per tensor, and a single flipped unit moves encoder.7.0.bias[251] by 1.9e-2 of the tensor's scale).  The test is
    therefore element-robust: per tensor the 99.5th percentile of |error| / max|grad| must be <= 5e-3 (about twice the fp32
    oracle's own distance from fp64; measured 2.2e-3 with fp32 kernels, 4.2e-3 with f16x3 forward + tf32x3 gradients), and no
    element may be off by more than 5 % of the tensor's scale."""
    import os
    from cleanumamba_b200.network import Net
    sums = json.load(open(os.path.join(__import__("conftest").GOLDEN, "full_init_seed0_sums.json")))["DNS-CleanUMamba-3N-E8"]
    torch.manual_seed(0)
    net = Net("CleanUMamba", dict(sums["config"], math_mode=math))
    sd = {k: v.detach().clone().double().requires_grad_() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    clean, noisy = orc.synth_batch(1, 1.0, seed=41)
    out_ref = orc.forward(sd, noisy.double(), differentiable=True, dtype=torch.float64)
    loss_ref = F.l1_loss(out_ref, clean.double()) + (out_ref ** 2).mean()
    loss_ref.backward()
    out = net(noisy.clone().cuda())
    loss = F.l1_loss(out, clean.cuda()) + (out ** 2).mean()
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    tol = 5e-3
    worst_q, worst_max, worst_k = 0.0, 0.0, None
    for k, p in net.named_parameters():
        gr = sd[k].grad
        assert p.grad is not None and p.grad.shape == gr.shape, k
        scale = gr.abs().max().item()
        if scale > 1e-12:
            e = ((p.grad.double().cpu() - gr).abs() / scale).flatten()
            q = torch.quantile(e, 0.995).item() if e.numel() > 200 else e.max().item()
            if q > worst_q:
                worst_q, worst_k = q, k
            worst_max = max(worst_max, e.max().item())
            assert q < tol, f"{k}: 99.5th percentile relative error {q:.3e}"
            assert e.max().item() < 5e-2, f"{k}: max relative error {e.max().item():.3e}"
    print(f"\n[E8-full grad {math}] worst per-tensor 99.5th-percentile relative gradient error {worst_q:.3e} ({worst_k}); "
          f"worst single element {worst_max:.3e}")


def test_gradients_accumulate_over_two_backwards_without_zero_grad():
    """Gradient accumulation (the reference's pruning loop, src/training/pruning.py:124-127) and ``zero_grad(set_to_none=False)``:
    no gradient handed to autograd may alias the engine's persistent flat gradient buffer.  Uses a model whose d_model and
    channel counts are multiples of 8 (no padding is cut away: the case where a ``.contiguous()`` would be a view)."""
    from cleanumamba_b200.network import Net
    cfg = dict(channels_input=1, channels_output=1, channels_H=16, max_H=64, encoder_n_layers=3, kernel_size=4, stride=2,
               tsfm_n_layers=2, tsfm_n_head=4, tsfm_d_model=64, tsfm_d_inner=128, math_mode="fp32")
    torch.manual_seed(1)
    net = Net("CleanUMamba", cfg)
    sd = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(2, 1, 3000, generator=g) * 0.1, torch.randn(2, 1, 3000, generator=g) * 0.3]
    for x in xs:            # oracle: two backward passes accumulate into .grad
        (orc.forward(sd, x, differentiable=True) ** 2).mean().backward()
    for x in xs:
        (net(x.clone().cuda()) ** 2).mean().backward()
    for k, p in net.named_parameters():
        gr = sd[k].grad
        scale = gr.abs().max().item()
        if scale > 1e-12:
            assert (p.grad.cpu() - gr).abs().max().item() / scale < 2e-3, k
    # set_to_none=False keeps the .grad tensors: the next backward must add to zeros, not to the engine's buffer
    net.zero_grad(set_to_none=False)
    for v in sd.values():
        v.grad = None
    (orc.forward(sd, xs[0], differentiable=True) ** 2).mean().backward()
    (net(xs[0].clone().cuda()) ** 2).mean().backward()
    for k, p in net.named_parameters():
        gr = sd[k].grad
        scale = gr.abs().max().item()
        if scale > 1e-12:
            assert (p.grad.cpu() - gr).abs().max().item() / scale < 2e-3, k
