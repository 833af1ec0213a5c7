"""GPU parity of every C-ABI operator against the CPU oracle / plain PyTorch fp32 on the same seeded inputs.
Tolerances are written next to each check (fp32 arithmetic, different summation order only)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import cleanumamba_oracle as orc

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


GEMM_TOL = {"fp32": 1e-5, "tf32x3": 6e-5, "bf16x3": 1.5e-4, "f16x3": 6e-5,
            "tf32": 1e-2}          # single-pass TF32: offered, never default, outside the model tolerance   # x3 modes: split error per product + truncating TMEM accumulation


def rel_err(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("b,d,l,n,with_z,with_h0", [
    (2, 64, 37, 64, True, False), (1, 2048, 130, 64, True, True), (3, 40, 50, 8, True, True),
    (2, 128, 33, 16, False, False), (2, 96, 20, 24, True, False), (1, 8, 5, 8, True, True), (2, 24, 624, 64, True, False),
    # few tokens + carried state: the streaming state-update kernel (selective_scan_step_kernel)
    (2, 64, 2, 64, True, True), (1, 2048, 1, 64, True, True), (3, 48, 2, 64, False, True), (2, 32, 1, 64, True, False), (2, 64, 3, 64, True, True),
    # many streams x whole 64-channel blocks: the TMA-staged state-update kernel (selective_scan_step_bulk_kernel), 1 and 2 tokens
    (48, 1024, 1, 64, True, True), (41, 1024, 2, 64, False, True), (300, 128, 2, 64, True, True),
    (48, 1024, 5, 64, True, True), (40, 1024, 16, 64, True, True), (300, 128, 9, 64, False, True)])      # ... up to 16 tokens per call
def test_selective_scan_matches_oracle(b, d, l, n, with_z, with_h0):
    from cleanumamba_b200 import ops
    g = torch.Generator().manual_seed(b * 1000 + d + l + n)
    u = torch.randn(b, d, l, generator=g)
    delta = torch.randn(b, d, l, generator=g) * 0.5
    A = -torch.exp(torch.randn(d, n, generator=g) * 0.5 + 0.5)
    Bm, Cm = torch.randn(b, n, l, generator=g), torch.randn(b, n, l, generator=g)
    D = torch.randn(d, generator=g)
    z = torch.randn(b, d, l, generator=g) if with_z else None
    bias = torch.randn(d, generator=g) * 0.5 - 2.0
    h0 = torch.randn(b, d, n, generator=g) if with_h0 else None
    y_ref, h_ref = orc.selective_scan(u, delta, A, Bm, Cm, D, z, bias, True, h0=h0, return_last_state=True)
    cu = lambda t: None if t is None else t.to(dev())  # noqa: E731
    y, h = ops.selective_scan_fn(cu(u), cu(delta), cu(A), cu(Bm), cu(Cm), cu(D), cu(z), cu(bias), True,
                                 return_last_state=True, initial_state=cu(h0))
    assert y.shape == y_ref.shape and h.shape == h_ref.shape
    assert rel_err(y, y_ref) < 2e-5      # ex2.approx + fp32 reassociation over <= 624 steps
    assert rel_err(h, h_ref) < 2e-5


@pytest.mark.parametrize("b,d,l,n,with_h0", [(1, 256, 2000, 64, True), (2, 128, 999, 16, False), (1, 64, 1300, 8, True), (1, 320, 257, 24, False)])
def test_selective_scan_segment_parallel_small_batch(b, d, l, n, with_h0):
    """Small batch x long sequence: the scan runs segment-parallel (local scans from h = 0, carry pass, second scan from the
    true start states) -- same y / last state as the oracle and as the time-sequential kernel."""
    import ctypes as C
    from cleanumamba_b200 import _lib, ops
    g = torch.Generator().manual_seed(b * 1000 + d + l + n)
    u = torch.randn(b, d, l, generator=g)
    delta = torch.randn(b, d, l, generator=g) * 0.5
    A = -torch.exp(torch.randn(d, n, generator=g) * 0.5 + 0.5)
    Bm, Cm = torch.randn(b, n, l, generator=g), torch.randn(b, n, l, generator=g)
    D, z = torch.randn(d, generator=g), torch.randn(b, d, l, generator=g)
    bias = torch.randn(d, generator=g) * 0.5 - 2.0
    h0 = torch.randn(b, d, n, generator=g) if with_h0 else None
    y_ref, h_ref = orc.selective_scan(u, delta, A, Bm, Cm, D, z, bias, True, h0=h0, return_last_state=True)
    cu = lambda t: None if t is None else t.to(dev())  # noqa: E731
    args = (cu(u), cu(delta), cu(A), cu(Bm), cu(Cm), cu(D), cu(z), cu(bias), True)
    # the library must actually choose the segmented path for this shape
    s = _lib.ScanDesc()
    s.batch, s.len, s.d, s.n_state = b, l, d, n
    assert _lib.init(torch.device(dev())).cum_selective_scan_workspace_bytes(C.byref(s)) > 0
    y, h = ops.selective_scan_fn(*args, return_last_state=True, initial_state=cu(h0))
    y1, h1 = ops.selective_scan_fn(*args, return_last_state=True, initial_state=cu(h0), segment_parallel=False)
    assert rel_err(y, y_ref) < 3e-5 and rel_err(h, h_ref) < 3e-5
    assert rel_err(y, y1) < 2e-5 and rel_err(h, h1) < 2e-5


def test_selective_scan_chunk_carry_equals_one_shot():
    """Carried state: scanning two halves with h_out -> h0 equals one pass (streaming contract)."""
    from cleanumamba_b200 import ops
    g = torch.Generator().manual_seed(7)
    b, d, l, n = 2, 64, 48, 64
    mk = lambda *s: torch.randn(*s, generator=g).to(dev())  # noqa: E731
    u, delta, Bm, Cm, z = mk(b, d, l), mk(b, d, l) * 0.3, mk(b, n, l), mk(b, n, l), mk(b, d, l)
    A, D, bias = -torch.exp(mk(d, n) * 0.3), mk(d), mk(d) * 0.1 - 1
    y, h = ops.selective_scan_fn(u, delta, A, Bm, Cm, D, z, bias, True, return_last_state=True)
    k = 19
    y1, h1 = ops.selective_scan_fn(u[..., :k], delta[..., :k], A, Bm[..., :k], Cm[..., :k], D, z[..., :k], bias, True, return_last_state=True)
    y2, h2 = ops.selective_scan_fn(u[..., k:], delta[..., k:], A, Bm[..., k:], Cm[..., k:], D, z[..., k:], bias, True, return_last_state=True, initial_state=h1)
    assert torch.equal(torch.cat([y1, y2], -1), y) and torch.equal(h2, h)


@pytest.mark.parametrize("b,d,l", [(2, 2048, 70), (3, 48, 5), (1, 8, 2), (2, 130, 33)])
def test_causal_conv1d_silu(b, d, l):
    from cleanumamba_b200 import ops
    g = torch.Generator().manual_seed(d + l)
    x, w, bias = torch.randn(b, d, l, generator=g), torch.randn(d, 4, generator=g), torch.randn(d, generator=g)
    ref = F.silu(F.conv1d(x, w[:, None], bias, padding=3, groups=d)[..., :l])
    y = ops.causal_conv1d_fn(x.to(dev()), w.to(dev()), bias.to(dev()), "silu")
    assert rel_err(y, ref) < 2e-6
    # carried conv state == processing the concatenation
    st = torch.zeros(b, d, 3, device=dev())
    k = max(1, l // 3)
    ya = ops.causal_conv1d_fn(x[..., :k].to(dev()), w.to(dev()), bias.to(dev()), "silu", conv_state=st)
    yb = ops.causal_conv1d_fn(x[..., k:].to(dev()), w.to(dev()), bias.to(dev()), "silu", conv_state=st) if l > k else ya[..., :0]
    assert rel_err(torch.cat([ya, yb], -1), ref) < 2e-6
    tail = F.pad(x, (3, 0))[..., -3:]
    assert torch.allclose(st.cpu(), tail)


@pytest.mark.parametrize("rows,c", [(50, 512), (7, 114), (33, 64), (5, 1000), (9, 16)])
def test_layer_norm_residual(rows, c):
    from cleanumamba_b200 import ops
    g = torch.Generator().manual_seed(rows + c)
    h, r = torch.randn(rows, c, generator=g), torch.randn(rows, c, generator=g) * 3
    w, b = torch.randn(c, generator=g), torch.randn(c, generator=g)
    ref = F.layer_norm(h + r, (c,), w, b, 1e-5)
    y, res = ops.layer_norm_residual(h.to(dev()), r.to(dev()), w.to(dev()), b.to(dev()), 1e-5)
    assert torch.allclose(res.cpu(), h + r, atol=0, rtol=0)
    assert (y.cpu() - ref).abs().max().item() < 5e-6
    y0, res0 = ops.layer_norm_residual(h.to(dev()), None, w.to(dev()), b.to(dev()), 1e-5)
    assert (y0.cpu() - F.layer_norm(h, (c,), w, b, 1e-5)).abs().max().item() < 5e-6 and torch.equal(res0.cpu(), h)


def _cl(x):  # (b, c, l) -> (b, l, c) contiguous on the GPU
    return x.permute(0, 2, 1).contiguous().to(dev())


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "bf16x3", "f16x3", "tf32"])
@pytest.mark.parametrize("b,cin,cout,l", [(2, 64, 128, 300), (1, 768, 768, 150), (3, 56, 72, 38), (1, 8, 8, 6), (2, 104, 200, 1030)])
def test_gemm_pointwise_and_glu(math, b, cin, cout, l):
    from cleanumamba_b200 import _lib, ops
    g = torch.Generator().manual_seed(cin + cout + l)
    x = torch.randn(b, cin, l, generator=g)
    w, bias = torch.randn(cout, cin, generator=g) / cin ** 0.5, torch.randn(cout, generator=g)
    ref = F.conv1d(x, w[:, :, None], bias)
    y = ops.gemm_bias_act(_cl(x), w[None].contiguous().to(dev()), bias.to(dev()), _lib.EPI_NONE, math=math)
    assert rel_err(y.permute(0, 2, 1), ref) < GEMM_TOL[math]
    yr = ops.gemm_bias_act(_cl(x), w[None].contiguous().to(dev()), bias.to(dev()), _lib.EPI_RELU, math=math)
    assert rel_err(yr.permute(0, 2, 1), F.relu(ref)) < GEMM_TOL[math]
    # GLU: interleave rows (a_c, b_c)
    H = cout // 2
    wi = torch.stack([w[:H], w[H:]], 1).reshape(cout, cin)
    bi = torch.stack([bias[:H], bias[H:]], 1).reshape(cout)
    add = torch.randn(b, l, H, generator=g)
    yg = ops.gemm_bias_act(_cl(x), wi[None].contiguous().to(dev()), bi.to(dev()), _lib.EPI_GLU["Sigmoid"],
                           addend=add.to(dev()), math=math)
    refg = orc.glu(ref) + add.permute(0, 2, 1)
    assert rel_err(yg.permute(0, 2, 1), refg) < GEMM_TOL[math]


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "bf16x3", "f16x3"])
@pytest.mark.parametrize("b,cin,cout,lout", [(2, 64, 128, 200), (1, 256, 512, 77), (2, 56, 40, 129), (1, 768, 768, 130)])
def test_gemm_as_strided_conv_and_transposed_conv(math, b, cin, cout, lout):
    """Conv1d(k=4,s=2) and ConvTranspose1d(k=4,s=2) expressed as 2-tap GEMMs (engine.py layouts)."""
    from cleanumamba_b200 import _lib, ops
    g = torch.Generator().manual_seed(cin * 3 + cout + lout)
    lin = 2 * lout + 2
    x = torch.randn(b, cin, lin, generator=g)
    w, bias = torch.randn(cout, cin, 4, generator=g) / (4 * cin) ** 0.5, torch.randn(cout, generator=g)
    ref = F.relu(F.conv1d(x, w, bias, stride=2))
    wt = torch.zeros(2, cout, 2 * cin)
    for s in range(2):
        for j in range(2):
            wt[s, :, j * cin:(j + 1) * cin] = w[:, :, 2 * s + j]
    a = _cl(x).view(b, lin // 2, 2 * cin)
    y = ops.gemm_bias_act(a, wt.to(dev()), bias.to(dev()), _lib.EPI_RELU, shifts=(0, 1), m=lout, math=math)
    assert rel_err(y.permute(0, 2, 1), ref) < GEMM_TOL[math]
    # transposed conv: (b, cin, lout) -> (b, cout, 2 lout + 2), + skip addend, ReLU before the add
    xt = torch.randn(b, cin, lout, generator=g)
    wT, bT = torch.randn(cin, cout, 4, generator=g) / (2 * cin) ** 0.5, torch.randn(cout, generator=g)
    skip = torch.randn(b, cout, lin, generator=g)
    refT = F.relu(F.conv_transpose1d(xt, wT, bT, stride=2)) + skip
    wp = torch.zeros(2, 2 * cout, cin)
    for s in range(2):
        for par in range(2):
            wp[s, par * cout:(par + 1) * cout] = wT[:, :, 2 * s + par].t()
    bp = torch.cat([bT, bT])
    add = _cl(skip).view(b, lout + 1, 2 * cout)
    yT = ops.gemm_bias_act(_cl(xt), wp.to(dev()), bp.to(dev()), _lib.EPI_RELU, shifts=(0, -1), m=lout + 1,
                           addend=add, math=math)
    assert rel_err(yT.reshape(b, lin, cout).permute(0, 2, 1), refT) < GEMM_TOL[math]


def test_wave_ends():
    """normalise (in place), conv_in (Cin=1 + pad), convt_out (Cout=1 + crop + scale)."""
    from cleanumamba_b200 import _lib
    lib = _lib.init(dev())
    g = torch.Generator().manual_seed(5)
    B, L, H = 3, 1000, 56
    x = torch.randn(B, L, generator=g) * 0.3 + 0.05
    xs = x.to(dev())
    std = torch.empty(B, device=dev())
    _lib.check(lib.cum_wave_normalize_fwd(xs.data_ptr(), std.data_ptr(), B, L, _lib.stream_ptr()), "norm")
    std_ref = x.std(dim=1) + 1e-3
    assert torch.allclose(std.cpu(), std_ref, rtol=1e-6, atol=0)
    assert torch.allclose(xs.cpu(), x / std_ref[:, None], rtol=2e-6, atol=1e-7)
    # conv_in with implicit right padding to 1022 -> rows_out 510
    Lp = orc.valid_length(L, 3)
    rows = (Lp - 4) // 2 + 1
    for H in (56, 64):      # 64 channels: the specialised kernel of the shipped geometry
        w, bias = torch.randn(H, 1, 4, generator=g), torch.randn(H, generator=g)
        ref = F.relu(F.conv1d(F.pad(x, (0, Lp - L))[:, None], w, bias, stride=2))
        y = torch.empty(B, rows, H, device=dev())
        xd, wd, bd = x.to(dev()), w[:, 0].t().contiguous().to(dev()), bias.to(dev())   # keep alive: raw pointers cross the ABI
        _lib.check(lib.cum_conv_in_fwd(xd.data_ptr(), L, B, L, wd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, H, 4, 2, 0, 0, 0,
                                       _lib.stream_ptr()), "conv_in")
        assert rel_err(y.permute(0, 2, 1), ref) < 1e-6
    # convt_out
    for Hc in (56, 64, 128):
        gin = torch.randn(B, Hc, rows, generator=g)
        wT, bT = torch.randn(Hc, 1, 4, generator=g), torch.randn(1, generator=g)
        scale = torch.rand(B, generator=g) + 0.5
        refo = F.conv_transpose1d(gin, wT, bT, stride=2)[..., :L] * scale[:, None, None]
        out = torch.empty(B, 1, L, device=dev())
        gcl, wd, sd = gin.permute(0, 2, 1).contiguous().to(dev()), wT[:, 0].t().contiguous().to(dev()), scale.to(dev())
        _lib.check(lib.cum_convt_out_fwd(gcl.data_ptr(), B, rows, Hc, wd.data_ptr(), float(bT), sd.data_ptr(), L,
                                         out.data_ptr(), L, 0, L, 4, 2, _lib.stream_ptr()), "convt_out")
        assert rel_err(out, refo) < 2e-6
        # `first` > 0 (streaming): same outputs, shifted window, per-group scale
        out2 = torch.empty(B, 1, L - 7, device=dev())
        sg = torch.rand(B, (L - 7 + 255) // 256, generator=g) + 0.5
        sgd = sg.to(dev())
        _lib.check(lib.cum_convt_out_fwd(gcl.data_ptr(), B, rows, Hc, wd.data_ptr(), float(bT), sgd.data_ptr(), 256,
                                         out2.data_ptr(), L - 7, 7, L - 7, 4, 2, _lib.stream_ptr()), "convt_out")
        full = F.conv_transpose1d(gin, wT, bT, stride=2)[..., 7:L]
        want = full * sg.repeat_interleave(256, 1)[:, None, : L - 7]
        assert rel_err(out2, want) < 2e-6


def test_bad_arguments_fail_loudly():
    from cleanumamba_b200 import _lib, ops
    a = torch.randn(1, 8, 6, device=dev())      # K = 6 is not a multiple of 4
    w = torch.randn(1, 8, 6, device=dev())
    with pytest.raises(RuntimeError, match="multiples of 4"):
        ops.gemm_bias_act(a, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.selective_scan_fn(*[torch.zeros(1, 4, 4)] * 2, torch.zeros(4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))


@pytest.mark.parametrize("b,cin,cout,l", [(2, 64, 128, 300), (1, 768, 768, 150), (2, 104, 200, 1030)])
def test_gemm_bf16_storage_mode(b, cin, cout, l):
    """CUM_MATH_BF16: bf16 activations / weights from HBM, fp32 accumulate, bf16 or fp32 output (reduced-precision variant)."""
    from cleanumamba_b200 import _lib
    lib = _lib.init(dev())
    g = torch.Generator().manual_seed(cin + l)
    x = torch.randn(b, l, cin, generator=g).to(dev()).to(torch.bfloat16).contiguous()
    w = (torch.randn(1, cout, cin, generator=g) / cin ** 0.5).to(dev()).to(torch.bfloat16).contiguous()
    bias = torch.randn(cout, generator=g).to(dev())
    ref = torch.relu(x.float() @ w[0].float().t() + bias)
    for out_dt in (torch.float32, torch.bfloat16):
        out = torch.empty(b, l, cout, dtype=out_dt, device=dev())
        d = _lib.GemmDesc()
        d.a, d.a_batch_stride, d.a_row_stride, d.a_rows, d.k, d.taps = x.data_ptr(), l * cin, cin, l, cin, 1
        d.w, d.ldw, d.bias = w.data_ptr(), cin, bias.data_ptr()
        d.c, d.c_batch_stride, d.c_row_stride, d.m, d.n, d.batch, d.epilogue = out.data_ptr(), l * cout, cout, l, cout, b, _lib.EPI_RELU
        d.math, d.out_bf16 = _lib.MATH_BF16, int(out_dt == torch.bfloat16)
        _lib.check(lib.cum_gemm_bias_act_fwd(C.byref(d), _lib.stream_ptr()), "gemm bf16")
        assert rel_err(out.float(), ref) < (2e-5 if out_dt == torch.float32 else 6e-3), out_dt


@pytest.mark.parametrize("math", ["f16x3", "tf32x3", "bf16x3", "tf32"])
@pytest.mark.parametrize("b,k,n,m,taps", [(1, 64, 256, 129, 1), (2, 128, 768, 700, 1), (1, 96, 160, 300, 1), (3, 256, 512, 257, 2),
                                          (1, 1536, 768, 1000, 2), (2, 64, 264, 256, 1), (1, 512, 1536, 5000, 1)])
def test_gemm_cta_pair_is_bit_identical_to_single_cta(math, b, k, n, m, taps):
    """Tiles wider than 128 columns run on CTA pairs (tcgen05 cta_group::2, 256-row tiles, half a weight tile per CTA).  Same
    products, same accumulation order: the result must equal the single-CTA kernel's bit for bit -- ragged row counts (a
    peer CTA with no rows), ragged column counts (N split unevenly over the pair), taps, batches, GLU + addend."""
    from cleanumamba_b200 import _lib, ops
    g = torch.Generator().manual_seed(k + n + m)
    rows = m + 1 if taps == 2 else m
    a = torch.randn(b, rows, k, generator=g).to(dev())
    w = (torch.randn(taps, n, k, generator=g) / (taps * k) ** 0.5).to(dev())
    bias = torch.randn(n, generator=g).to(dev())
    add = torch.randn(b, m, n // 2, generator=g).to(dev())
    shifts = (0, 1) if taps == 2 else (0, 0)
    for epi, ad in ((_lib.EPI_RELU, None), (_lib.EPI_GLU["Sigmoid"], add)):
        pair = ops.gemm_bias_act(a, w, bias, epi, shifts=shifts, m=m, addend=ad, math=math, cta_pair=1)
        single = ops.gemm_bias_act(a, w, bias, epi, shifts=shifts, m=m, addend=ad, math=math, cta_pair=-1)
        assert torch.equal(pair, single), (pair - single).abs().max().item()
    ref = torch.relu(sum(F.pad(a, (0, 0, 0, 1))[:, s:s + m].double().cpu() @ w[t].double().cpu().t()
                         for t, s in enumerate(shifts[:taps])) + bias.double().cpu())
    assert rel_err(ops.gemm_bias_act(a, w, bias, _lib.EPI_RELU, shifts=shifts, m=m, math=math), ref) < GEMM_TOL[math]


@pytest.mark.parametrize("cta_pair", [0, -1])
@pytest.mark.parametrize("b,k,n,m,taps", [(1, 64, 128, 300, 1), (2, 128, 768, 700, 1), (1, 96, 160, 300, 1), (3, 256, 512, 257, 2),
                                          (1, 1536, 768, 1000, 2), (2, 64, 264, 256, 1), (2, 56, 72, 77, 1)])
def test_gemm_hl16_activation_planes(b, k, n, m, taps, cta_pair):
    """f16x3 with activations stored as fp16 hi/lo planes (cum_gemm_desc.a_lo / c_lo / addend_lo): the operands that reach the
    tensor cores are the ones the in-kernel splitter would have produced, so (1) pre-split input + fp32 output is bit-identical
    to the fp32-input kernel, (2) a split output is exactly the hi/lo split of that fp32 result, also with a split addend."""
    from cleanumamba_b200 import _lib, ops
    g = torch.Generator().manual_seed(k * 7 + n + m)
    rows = m + 1 if taps == 2 else m
    a = (torch.randn(b, rows, k, generator=g) * torch.exp(torch.randn(b, rows, 1, generator=g) * 3)).to(dev())   # wide dynamic range
    w = (torch.randn(taps, n, k, generator=g) / (taps * k) ** 0.5).to(dev())
    bias = torch.randn(n, generator=g).to(dev())
    shifts = (0, 1) if taps == 2 else (0, 0)
    ah = ops.split_hl16(a)
    for epi in (_lib.EPI_RELU, _lib.EPI_GLU["Sigmoid"]):
        n_out = n // 2 if epi >= 8 else n
        ref = ops.gemm_bias_act(a, w, bias, epi, shifts=shifts, m=m, math="f16x3", cta_pair=cta_pair)
        got = ops.gemm_bias_act(ah, w, bias, epi, shifts=shifts, m=m, math="f16x3", cta_pair=cta_pair)
        assert torch.equal(got, ref), (got - ref).abs().max().item()
        add = torch.randn(b, m, n_out, generator=g).to(dev())
        addh = ops.split_hl16(add)
        add_q = addh[0].float() + addh[1].float()
        ref2 = ops.gemm_bias_act(a, w, bias, epi, shifts=shifts, m=m, addend=add_q, math="f16x3", cta_pair=cta_pair)
        out = ops.gemm_bias_act(ah, w, bias, epi, shifts=shifts, m=m, addend=addh, math="f16x3", cta_pair=cta_pair, out_hl16=True)
        assert torch.equal(out, ops.split_hl16(ref2))
        # the addend's format is independent of the output's (a skip tensor keeps the format its encoder layer wrote)
        assert torch.equal(ops.gemm_bias_act(a, w, bias, epi, shifts=shifts, m=m, addend=addh, math="f16x3", cta_pair=cta_pair), ref2)
        assert torch.equal(ops.gemm_bias_act(ah, w, bias, epi, shifts=shifts, m=m, addend=add_q, math="f16x3", cta_pair=cta_pair,
                                             out_hl16=True), ops.split_hl16(ref2))


def test_conv_in_hl16_matches_fp32_kernel():
    from cleanumamba_b200 import _lib, ops
    lib = _lib.init(dev())
    g = torch.Generator().manual_seed(11)
    B, L = 3, 1000
    x = (torch.randn(B, L, generator=g) * 0.3).to(dev())
    rows = (L - 4) // 2 + 1
    for H in (56, 64):
        w, bias = torch.randn(4, H, generator=g).to(dev()), torch.randn(H, generator=g).to(dev())
        y = torch.empty(B, rows, H, device=dev())
        _lib.check(lib.cum_conv_in_fwd(x.data_ptr(), L, B, L, w.data_ptr(), bias.data_ptr(), y.data_ptr(), rows, H, 4, 2, 0, 0, 0,
                                       _lib.stream_ptr()), "conv_in")
        yh = torch.empty(2, B, rows, H, dtype=torch.float16, device=dev())
        _lib.check(lib.cum_conv_in_hl16_fwd(x.data_ptr(), L, B, L, w.data_ptr(), bias.data_ptr(), yh[0].data_ptr(), yh[1].data_ptr(),
                                            rows, H, 4, 2, _lib.stream_ptr()), "conv_in_hl16")
        assert torch.equal(yh, ops.split_hl16(y))


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "bf16x3", "f16x3", "tf32"])
@pytest.mark.parametrize("gate", ["ReLU", "SiLU", "GELU"])
def test_gemm_generic_glu_gates_and_unary_silu(gate, math):
    """CUM_EPI_GLU_{RELU,SILU,GELU} (layers.py:17-24) and the unary CUM_EPI_SILU epilogue against PyTorch in every math mode."""
    from cleanumamba_b200 import _lib, ops
    b, cin, cout, l = 2, 96, 144, 333
    g = torch.Generator().manual_seed(17)
    x = torch.randn(b, cin, l, generator=g)
    w, bias = torch.randn(cout, cin, generator=g) / cin ** 0.5, torch.randn(cout, generator=g)
    ref = F.conv1d(x, w[:, :, None], bias)
    H = cout // 2
    wi = torch.stack([w[:H], w[H:]], 1).reshape(cout, cin)
    bi = torch.stack([bias[:H], bias[H:]], 1).reshape(cout)
    add = torch.randn(b, l, H, generator=g)
    yg = ops.gemm_bias_act(_cl(x), wi[None].contiguous().to(dev()), bi.to(dev()), _lib.EPI_GLU[gate], addend=add.to(dev()), math=math)
    refg = orc.glu(ref, gate) + add.permute(0, 2, 1)
    assert rel_err(yg.permute(0, 2, 1), refg) < GEMM_TOL[math]
    ys = ops.gemm_bias_act(_cl(x), w[None].contiguous().to(dev()), bias.to(dev()), _lib.EPI_SILU, math=math)
    assert rel_err(ys.permute(0, 2, 1), F.silu(ref)) < GEMM_TOL[math]


def test_library_caches_survive_shutdown():
    """cum_shutdown() drops the tensor-map cache / per-device kernel attributes; the library keeps working afterwards."""
    from cleanumamba_b200 import _lib, ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 64, 200, generator=g)
    w = torch.randn(128, 64, generator=g) / 8
    want = F.conv1d(x, w[:, :, None])
    for _ in range(2):
        y = ops.gemm_bias_act(_cl(x), w[None].contiguous().to(dev()), None, _lib.EPI_NONE, math="f16x3")
        assert rel_err(y.permute(0, 2, 1), want) < GEMM_TOL["f16x3"]
        assert _lib.load().cum_shutdown() == 0


def _split_f16(w_dev):
    """fp16 hi / lo halves of 2^k w (k so that max|2^k w| is in [8, 16)) like engine._pack; returns hi, lo, 1 / 2^k."""
    import math
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(dev()))
    a = w_dev.abs().max().item()
    e = max(-14, min(14, int(math.floor(math.log2(8.0 / a))))) if a > 0 else 0
    hi, lo = torch.empty_like(w_dev, dtype=torch.float16), torch.empty_like(w_dev, dtype=torch.float16)
    _lib.check(lib.cum_split_f16(w_dev.data_ptr(), hi.data_ptr(), lo.data_ptr(), w_dev.numel(), float(2.0 ** e), _lib.stream_ptr()), "split")
    return hi, lo, float(2.0 ** -e)


@pytest.mark.parametrize("hc,ho", [(64, 64), (56, 56), (32, 32), (24, 40)])
@pytest.mark.parametrize("b,length", [(1, 200), (2, 4099), (3, 33000), (1, 6), (4, 160000)])       # last: > 16 tiles per CTA
def test_fused_enc0_block(b, length, hc, ho):
    """cum_enc0_block_fwd == F.pad + Conv1d(1,Hc,4,2) + ReLU + Conv1d(Hc,2 Ho,1) + GLU (CleanUMamba.py:108-113), ragged lengths,
    the shipped 64-channel geometry and narrower (pruned) widths that run zero-padded inside the 64 x 128 tile."""
    import ctypes as C
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(dev()))
    g = torch.Generator().manual_seed(length + hc)
    x = torch.randn(b, 1, length, generator=g)
    w0, b0 = torch.randn(hc, 1, 4, generator=g) * 0.5, torch.randn(hc, generator=g) * 0.1
    w1, b1 = torch.randn(2 * ho, hc, generator=g) / 8, torch.randn(2 * ho, generator=g) * 0.1
    padded = max(length, 4) + (max(length, 4) % 2)
    rows = (padded - 4) // 2 + 1
    ref = orc.glu(F.conv1d(F.relu(F.conv1d(F.pad(x, (0, padded - length)), w0, b0, stride=2)), w1[:, :, None], b1))      # (b, ho, rows)
    wi = torch.stack([w1[:ho], w1[ho:]], 1).reshape(2 * ho, hc).contiguous().to(dev())
    bi = torch.stack([b1[:ho], b1[ho:]], 1).reshape(2 * ho).contiguous().to(dev())
    hi, lo, inv = _split_f16(wi)
    xd = x[:, 0].contiguous().to(dev())
    cw, cb = w0[:, 0, :].t().contiguous().to(dev()), b0.to(dev())
    out = torch.full((b, rows, ho), float("nan"), device=dev())
    d = _lib.Enc0BlockDesc()
    d.x, d.x_stride, d.batch, d.length = xd.data_ptr(), length, b, length
    d.conv_w, d.conv_b, d.glu_w_hi, d.glu_w_lo, d.glu_b = cw.data_ptr(), cb.data_ptr(), hi.data_ptr(), lo.data_ptr(), bi.data_ptr()
    d.acc_scale, d.w_lo_is_zero, d.out, d.rows_out, d.channels, d.channels_out = inv, 0, out.data_ptr(), rows, hc, ho
    _lib.check(lib.cum_enc0_block_fwd(C.byref(d), _lib.stream_ptr()), "enc0_block")
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    assert rel_err(out.permute(0, 2, 1), ref) < GEMM_TOL["f16x3"]
    d.channels = 72
    assert lib.cum_enc0_block_fwd(C.byref(d), _lib.stream_ptr()) != 0      # wider than the 64-channel tile: not served


@pytest.mark.parametrize("cin,hg", [(64, 64), (56, 64), (32, 32), (40, 24)])
@pytest.mark.parametrize("b,rows,crop", [(1, 100, 0), (2, 2047, 130), (3, 16500, 0), (1, 1, 1), (4, 80126, 254)])   # last: > 16 tiles per CTA
def test_fused_dec_last_block(b, rows, crop, cin, hg):
    """cum_dec_last_block_fwd == Conv1d(Cin,2 Hg,1) + GLU + ConvTranspose1d(Hg,1,4,2) + crop + * std (CleanUMamba.py:121-128, :318-319)."""
    import ctypes as C
    from cleanumamba_b200 import _lib
    lib = _lib.init(torch.device(dev()))
    g = torch.Generator().manual_seed(rows + cin)
    a = torch.randn(b, cin, rows, generator=g)
    w1, b1 = torch.randn(2 * hg, cin, generator=g) / 8, torch.randn(2 * hg, generator=g) * 0.1
    wt, bt = torch.randn(hg, 1, 4, generator=g) / 8, torch.randn(1, generator=g) * 0.1
    std = torch.rand(b, generator=g) + 0.5
    full = F.conv_transpose1d(orc.glu(F.conv1d(a, w1[:, :, None], b1)), wt, bt, stride=2)      # (b, 1, 2 rows + 2)
    length = 2 * rows + 2 - crop
    ref = full[:, :, :length] * std[:, None, None]
    wi = torch.stack([w1[:hg], w1[hg:]], 1).reshape(2 * hg, cin).contiguous().to(dev())
    bi = torch.stack([b1[:hg], b1[hg:]], 1).reshape(2 * hg).contiguous().to(dev())
    hi, lo, inv = _split_f16(wi)
    acl = _cl(a)
    tw = wt[:, 0, :].t().contiguous().to(dev())
    sd = std.to(dev())
    out = torch.full((b, 1, length), float("nan"), device=dev())
    d = _lib.DecLastBlockDesc()
    d.a, d.batch, d.rows_in = acl.data_ptr(), b, rows
    d.glu_w_hi, d.glu_w_lo, d.glu_b, d.acc_scale, d.w_lo_is_zero = hi.data_ptr(), lo.data_ptr(), bi.data_ptr(), inv, 0
    d.convt_w, d.convt_bias, d.scale = tw.data_ptr(), float(bt[0]), sd.data_ptr()
    d.out, d.out_stride, d.out_length, d.channels, d.channels_gated = out.data_ptr(), length, length, cin, hg
    _lib.check(lib.cum_dec_last_block_fwd(C.byref(d), _lib.stream_ptr()), "dec_last_block")
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    assert rel_err(out, ref) < GEMM_TOL["f16x3"]


@pytest.mark.parametrize("shape,dim,off,n_ch,heads", [((16, 12, 4), 0, 0, 16, 1), ((16, 12, 4), 1, 0, 12, 1), ((4096, 512), 0, 2048, 512, 4),
                                                      ((160, 2048), 1, 5, 2000, 1), ((768,), 0, 0, 768, 1), ((53, 111, 4), 1, 3, 50, 2),
                                                      ((2048, 64), 1, 0, 64, 1)])
def test_channel_importance_kernel_matches_oracle(shape, dim, off, n_ch, heads):
    """cum_channel_importance_fwd + importance.channel_importances == the reference's PruningModule.channel_importances
    (pruninggroup.py:160-226; restated in the oracle, which is pinned to the live reference on CPU)."""
    from cleanumamba_b200 import importance
    g = torch.Generator().manual_seed(sum(shape) + dim)
    w, gr = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g) * 1e-3
    want = orc.channel_importances(w, gr, dim, off, n_ch, heads)
    got = importance.channel_importances(w.to(dev()), gr.to(dev()), dim, off, n_ch, heads)
    for k in ("weight", "grad", "taylor_individual", "taylor_squared_individual"):
        assert rel_err(got[k], want[k]) < 1e-5, k
    # |sum w g| cancels: bound the error by the un-cancelled magnitude
    assert ((got["taylor_group"].cpu() - want["taylor_group"]).abs() <= 1e-5 * want["taylor_individual"] + 1e-12).all()
    assert got["n_parameters"] == want["n_parameters"]
    only_w = importance.channel_importances(w.to(dev()), None, dim, off, n_ch, heads)
    assert only_w["grad"] is None and rel_err(only_w["weight"], want["weight"]) < 1e-5


@pytest.mark.parametrize("b,t", [(2, 16000), (3, 5003), (1, 1025)])
def test_fused_mr_stft_loss_matches_pytorch_restatement(b, t):
    """FusedMultiResolutionSTFTLoss (DFT as tcgen05 GEMM + fused reductions, own backward) against loss.MultiResolutionSTFTLoss
    -- the PyTorch restatement that tests/test_cpu_* pin to the reference's stft_loss.py run live -- values and d/dx."""
    from cleanumamba_b200.fused_loss import FusedMultiResolutionSTFTLoss
    from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG, MultiResolutionSTFTLoss
    g = torch.Generator().manual_seed(t)
    clean = torch.randn(b, t, generator=g) * 0.1
    pred = (clean + torch.randn(b, t, generator=g) * 0.03).requires_grad_()
    sc_ref, mag_ref = MultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)(pred, clean)
    (2.0 * sc_ref + 0.7 * mag_ref).backward()
    pd = pred.detach().to(dev()).requires_grad_()
    sc, mag = FusedMultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)(pd, clean.to(dev()))
    (2.0 * sc + 0.7 * mag).backward()
    assert abs(sc.item() - sc_ref.item()) < 1e-5 * abs(sc_ref.item()) + 1e-7
    assert abs(mag.item() - mag_ref.item()) < 1e-5 * abs(mag_ref.item()) + 1e-7
    scale = pred.grad.abs().max().item()
    err = (pd.grad.cpu() - pred.grad).abs().max().item()
    print(f"\n[fused MR-STFT loss b={b} t={t}] sc {sc.item():.6f} mag {mag.item():.6f} grad max-abs err / scale {err / scale:.2e}")
    assert err / scale < 1e-3


@pytest.mark.parametrize("name", ["e8_pruned_500k", "mini_mamba_442k"])
def test_standalone_submodule_forwards_match_oracle(name):
    """The module tree's submodules are individually callable like the reference's (pruning / analysis code does that):
    ``Block.forward`` / ``Mamba.forward`` (mamba_ssm slow path) and the GLU ``Activation`` on the library's kernels, against the
    oracle's mixer / LayerNorm / GLU on a trained checkpoint."""
    import json
    from conftest import load_golden
    from cleanumamba_b200.network import Net
    fx = load_golden(name)
    net = Net("CleanUMamba", json.loads(fx["config"]))
    net.load_pruned_state_dict(fx["state_dict"])
    net = net.cuda().float().eval()
    sd = {k: v.float() for k, v in fx["state_dict"].items()}
    dm = net.tsfm_conv1.weight.shape[0]
    g = torch.Generator().manual_seed(5)
    h = torch.randn(2, 37, dm, generator=g)
    res = torch.randn(2, 37, dm, generator=g)
    blk = net.tsfm_Mamba_layers[1]
    p = "tsfm_Mamba_layers.1."
    want_res = h + res
    normed = F.layer_norm(want_res, (dm,), sd[p + "norm.weight"], sd[p + "norm.bias"], blk.norm.eps)
    want = orc.mamba_mixer(normed, sd, p + "mixer.")
    out, r = blk(h.cuda(), res.cuda())
    assert out.shape == want.shape and torch.equal(r.cpu(), want_res)
    assert rel_err(out, want) < 2e-4                       # f16x3 products through four GEMMs + the scan
    out0 = blk.mixer(normed.cuda())
    assert rel_err(out0, want) < 2e-4
    with pytest.raises(NotImplementedError):
        blk.mixer(normed.cuda(), inference_params=object())
    x = torch.randn(2, 2 * 24, 50, generator=g)
    y = net.encoder[0][3](x.cuda())                         # layers.Activation (GLU)
    assert rel_err(y, orc.glu(x)) < 1e-6
