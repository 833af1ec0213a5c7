"""Operator-level drop-ins with the SIGNATURES AND LAYOUTS of the third-party operators the reference calls
(mamba-ssm==1.2.2 / causal-conv1d==1.1.0; SURVEY.md §8b "lower boundary"):

    selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False)
    causal_conv1d_fn(x, weight, bias=None, activation=None)
    layer_norm_residual(h, residual, weight, bias, eps)

They accept the reference's channel-major (b, d, l) tensors, transpose to the channels-last layout of the kernels,
and call the C ABI.  Inside ``CleanUMamba.forward`` the engine calls the same kernels without these transposes.
CUDA tensors only -- no fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ScanDesc, check, ptr

LOG2E = 1.4426950408889634


def _need_cuda(*ts):
    for t in ts:
        if t is not None and t.device.type != "cuda":
            raise RuntimeError("cleanumamba_b200.ops: CUDA tensors required (no CPU fallback)")


def _cl(t):
    """(b, c, l) -> contiguous channels-last (b, l, c) fp32."""
    return t.detach().float().permute(0, 2, 1).contiguous()


@torch.no_grad()
def scan_workspace(lib, desc, device):
    """Scratch for the segment-parallel scan of small batches (None when the problem runs time-sequentially); sets the
    descriptor's workspace fields.  The caller keeps the returned tensor alive until the call has been issued."""
    nbytes = lib.cum_selective_scan_workspace_bytes(C.byref(desc))
    if nbytes <= 0:
        return None
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
    desc.workspace, desc.workspace_bytes = ws.data_ptr(), nbytes
    return ws


def selective_scan_fn(u, delta, A, B, C_, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False, initial_state=None, segment_parallel=True):
    """u, delta, z: (b, d, l); A: (d, n); B, C: (b, n, l).  Returns y (b, d, l) [, last_state (b, d, n)].
    ``initial_state`` (b, d, n) is an extension (the dependency's kernel always starts from h = 0)."""
    _need_cuda(u, delta, A, B, C_, D, z, delta_bias, initial_state)
    lib = _lib.init(u.device)
    b, d, l = u.shape
    n = A.shape[1]
    ucl, dcl, Bcl, Ccl = _cl(u), _cl(delta), _cl(B), _cl(C_)
    zcl = _cl(z) if z is not None else None
    y = torch.empty(b, l, d, dtype=torch.float32, device=u.device)
    a2 = (A.detach().float() * LOG2E).contiguous()
    Df = D.detach().float().contiguous() if D is not None else None
    bias = delta_bias.detach().float().contiguous() if delta_bias is not None else None
    h0 = initial_state.detach().float().contiguous() if initial_state is not None else None
    h_out = torch.empty(b, d, n, dtype=torch.float32, device=u.device) if return_last_state else None
    s = ScanDesc()
    s.u, s.u_bs, s.u_rs = ucl.data_ptr(), l * d, d
    s.delta, s.dl_bs, s.dl_rs = dcl.data_ptr(), l * d, d
    s.z, s.z_bs, s.z_rs = ptr(zcl), l * d, d
    s.Bm, s.B_bs, s.B_rs = Bcl.data_ptr(), l * n, n
    s.Cm, s.C_bs, s.C_rs = Ccl.data_ptr(), l * n, n
    s.y, s.y_bs, s.y_rs = y.data_ptr(), l * d, d
    s.a2, s.Dskip, s.delta_bias, s.h0, s.h_out = a2.data_ptr(), ptr(Df), ptr(bias), ptr(h0), ptr(h_out)
    s.batch, s.len, s.d, s.n_state, s.delta_softplus = b, l, d, n, int(bool(delta_softplus))
    ws = scan_workspace(lib, s, u.device) if segment_parallel else None
    check(lib.cum_selective_scan_fwd(C.byref(s), _lib.stream_ptr()), "cum_selective_scan_fwd")
    out = y.permute(0, 2, 1).to(u.dtype)
    return (out, h_out) if return_last_state else out


@torch.no_grad()
def causal_conv1d_fn(x, weight, bias=None, activation=None, conv_state=None):
    """x: (b, d, l); weight: (d, width); bias: (d).  y = act(causal depthwise conv).  Only activation "silu"/"swish"
    is fused (what Mamba uses).  ``conv_state`` (b, d, width-1), if given, supplies the inputs before t = 0 and is
    updated in place to the last width-1 inputs (the role of causal_conv1d_update)."""
    if activation not in ("silu", "swish"):
        raise NotImplementedError("cleanumamba_b200.causal_conv1d_fn: activation must be 'silu' (Mamba's use)")
    _need_cuda(x, weight, bias, conv_state)
    lib = _lib.init(x.device)
    b, d, l = x.shape
    width = weight.shape[1]
    dp = (d + 3) // 4 * 4
    xcl = torch.zeros(b, l, dp, dtype=torch.float32, device=x.device)
    xcl[:, :, :d] = x.detach().float().permute(0, 2, 1)
    w = torch.zeros(width, dp, dtype=torch.float32, device=x.device)
    w[:, :d] = weight.detach().float().t()
    bv = torch.zeros(dp, dtype=torch.float32, device=x.device)
    if bias is not None:
        bv[:d] = bias.detach().float()
    y = torch.empty(b, l, dp, dtype=torch.float32, device=x.device)
    st = None
    if conv_state is not None:
        st = torch.zeros(b, width - 1, dp, dtype=torch.float32, device=x.device)
        st[:, :, :d] = conv_state.detach().float().permute(0, 2, 1)
    check(lib.cum_dwconv_silu_fwd(xcl.data_ptr(), l * dp, dp, w.data_ptr(), bv.data_ptr(), y.data_ptr(), ptr(st),
                                  ptr(st), b, l, dp, width, _lib.stream_ptr()), "cum_dwconv_silu_fwd")
    if conv_state is not None:
        conv_state.copy_(st[:, :, :d].permute(0, 2, 1))
    return y[:, :, :d].permute(0, 2, 1).to(x.dtype)


@torch.no_grad()
def layer_norm_residual(h, residual, weight, bias, eps=1e-5):
    """(rows..., c) tensors: returns (LayerNorm(h + residual), h + residual) -- Block.forward's add + norm."""
    _need_cuda(h, residual, weight, bias)
    lib = _lib.init(h.device)
    c = h.shape[-1]
    cp = (c + 3) // 4 * 4
    rows = h.numel() // c

    def padc(t):
        o = torch.zeros(rows, cp, dtype=torch.float32, device=h.device)
        o[:, :c] = t.detach().float().reshape(rows, c)
        return o
    hp = padc(h)
    rp = padc(residual) if residual is not None else None
    g = torch.zeros(cp, dtype=torch.float32, device=h.device)
    g[:c] = weight.detach().float()
    be = torch.zeros(cp, dtype=torch.float32, device=h.device)
    be[:c] = bias.detach().float()
    ro = torch.empty_like(hp)
    no = torch.empty_like(hp)
    check(lib.cum_ln_residual_fwd(hp.data_ptr(), ptr(rp), ro.data_ptr(), no.data_ptr(), g.data_ptr(), be.data_ptr(),
                                  float(eps), rows, c, cp, _lib.stream_ptr()), "cum_ln_residual_fwd")
    return no[:, :c].reshape(h.shape), ro[:, :c].reshape(h.shape)


def _linear(x, weight, bias=None, math="f16x3"):
    """(b, l, k) x (n, k)^T (+ bias) on the tap-GEMM; k and n padded to multiples of 8 with zeros."""
    b, l, k = x.shape
    n = weight.shape[0]
    kp, np_ = (k + 7) // 8 * 8, (n + 7) // 8 * 8
    a = x.detach().float()
    if kp != k or not a.is_contiguous():
        a = torch.nn.functional.pad(a, (0, kp - k)).contiguous()
    w = torch.zeros(1, np_, kp, dtype=torch.float32, device=x.device)
    w[0, :n, :k] = weight.detach().float()
    bv = None
    if bias is not None:
        bv = torch.zeros(np_, dtype=torch.float32, device=x.device)
        bv[:n] = bias.detach().float()
    return gemm_bias_act(a, w, bv, _lib.EPI_NONE, math=math)[..., :n]


@torch.no_grad()
def mamba_mixer_forward(mixer, hidden_states, inference_params=None):
    """Stand-alone ``Mamba.forward`` (mamba_ssm 1.2.2 ``Mamba.forward`` slow path) on the library's kernels:
    in_proj -> causal depthwise conv + SiLU -> x_proj -> dt_proj -> selective scan (softplus, D skip, SiLU(z) gate) -> out_proj.
    hidden_states: (batch, len, d_model) on CUDA.  Inference only; inside ``CleanUMamba.forward`` the engine runs the same kernels on
    its packed weights without the layout changes made here.  Step mode (``inference_params``) is served by ``stream_session``."""
    if inference_params is not None:
        raise NotImplementedError("cleanumamba_b200: stand-alone Mamba.step (inference_params) is served through stream_session()")
    _need_cuda(hidden_states)
    math = getattr(mixer, "math_mode", "f16x3")
    di, n = mixer.A_log.shape
    r = mixer.dt_proj.weight.shape[1]
    xz = _linear(hidden_states, mixer.in_proj.weight, mixer.in_proj.bias, math)                 # (b, l, 2 di)
    x, z = xz[..., :di].permute(0, 2, 1), xz[..., di:].permute(0, 2, 1)                          # (b, di, l)
    x = causal_conv1d_fn(x, mixer.conv1d.weight[:, 0, :], mixer.conv1d.bias, "silu")
    x_dbl = _linear(x.permute(0, 2, 1), mixer.x_proj.weight, None, math)                          # (b, l, r + 2 n)
    dt = _linear(x_dbl[..., :r], mixer.dt_proj.weight, None, math).permute(0, 2, 1)               # (b, di, l); bias goes in as delta_bias
    Bm, Cm = x_dbl[..., r: r + n].permute(0, 2, 1), x_dbl[..., r + n:].permute(0, 2, 1)          # (b, n, l)
    y = selective_scan_fn(x, dt, -torch.exp(mixer.A_log.detach().float()), Bm, Cm, mixer.D.detach().float(), z,
                          mixer.dt_proj.bias.detach().float(), True)
    return _linear(y.permute(0, 2, 1), mixer.out_proj.weight, mixer.out_proj.bias, math).to(hidden_states.dtype)


@torch.no_grad()
def block_forward(block, hidden_states, residual=None, inference_params=None):
    """Stand-alone ``Block.forward`` (mamba_ssm ``Block``: residual add -> LayerNorm -> mixer; returns (hidden_states, residual))."""
    _need_cuda(hidden_states, residual)
    normed, res = layer_norm_residual(hidden_states, residual, block.norm.weight, block.norm.bias, block.norm.eps)
    return block.mixer(normed.to(hidden_states.dtype), inference_params=inference_params), res.to(hidden_states.dtype)


@torch.no_grad()
def glu_forward(x, kind="Sigmoid"):
    """Stand-alone GLU of layers.Activation (/root/reference/src/network/layers.py:26-33, bypass_channels = 0): (b, 2c, l) ->
    a * sigmoid(b) on the library's gate kernel (the model's forward fuses the gate into the 1x1-conv GEMM epilogue instead)."""
    if kind != "Sigmoid":
        raise NotImplementedError("cleanumamba_b200: the stand-alone GLU kernel implements the Sigmoid gate (other gates: GEMM epilogues only)")
    _need_cuda(x)
    lib = _lib.init(x.device)
    b, c2, l = x.shape
    c = c2 // 2
    cp = (c + 3) // 4 * 4
    z = torch.zeros(b * l, 2 * cp, dtype=torch.float32, device=x.device)          # interleaved (a_c, b_c) rows, see cum_glu_fwd
    xt = x.detach().float().permute(0, 2, 1).reshape(b * l, c2)
    z[:, 0: 2 * c: 2] = xt[:, :c]
    z[:, 1: 2 * c: 2] = xt[:, c:]
    out = torch.empty(b * l, cp, dtype=torch.float32, device=x.device)
    check(lib.cum_glu_fwd(z.data_ptr(), None, out.data_ptr(), b * l, cp, _lib.stream_ptr()), "cum_glu_fwd")
    return out[:, :c].reshape(b, l, c).permute(0, 2, 1).to(x.dtype)


@torch.no_grad()
def split_hl16(x):
    """fp32 -> the two-plane fp16 "hl16" format (2, *x.shape): hi = fp16(x), lo = fp16(x - hi)  (see cum_gemm_desc.a_lo)."""
    hi = x.clamp(-65504.0, 65504.0).to(torch.float16)
    lo = (x - hi.float()).clamp(-65504.0, 65504.0).to(torch.float16)
    return torch.stack([hi, lo])


def gemm_bias_act(a, w, bias=None, epilogue=_lib.EPI_NONE, shifts=(0, 0), m=None, addend=None, math="fp32", cta_pair=0,
                  out_hl16=False, plane_major=None):
    """Raw tap-GEMM on channels-last tensors (thin wrapper of cum_gemm_bias_act_fwd, used by tests / benchmarks).
    a: (batch, rows, K) fp32 contiguous, K % 4 == 0;  w: (taps, N, K), N % 8 == 0;  bias: (N);  addend: (batch, m, N_out).
    out[b, i, :] = EPI(bias + sum_s W_s . a[b, i + shifts[s], :]) + addend[b, i, :]   (rows outside a read as 0).
    f16x3 only: ``a`` may be an hl16 tensor (2, batch, rows, K) float16 (``split_hl16``); ``out_hl16`` returns the result in
    that format, and the addend must then be hl16 as well.
    ``plane_major=dict(batch=, plane0=, step=, n_half=)``: ``a`` is a stack (planes, rows, plane_k) of planes (the time-major streaming
    FIFOs, cum_gemm_desc.a_planes): batch item b, tap s, K offset kk read plane plane0 + step * (b' + shifts[s]) + kk // plane_k,
    b' = b >> 1 with n_half (then w holds 2 n rows per tap, bias n entries, and item b uses rows (b & 1) n ...)."""
    _need_cuda(a, w, bias, addend)
    lib = _lib.init(a.device)
    a_planes = a if a.dtype == torch.float16 else None
    batch, rows, k = a.shape[-3:]
    taps, n, kw = w.shape
    pm = plane_major
    if pm is not None:
        n_planes, plane_k = batch, k
        batch, k = pm["batch"], kw
        if pm.get("n_half"):
            n //= 2
    assert kw == k and a.is_contiguous() and w.is_contiguous()
    m = rows if m is None else m
    n_out = n // 2 if epilogue >= 8 else n
    out = (torch.empty(2, batch, m, n_out, dtype=torch.float16, device=a.device) if out_hl16
           else torch.empty(batch, m, n_out, dtype=torch.float32, device=a.device))
    d = _lib.GemmDesc()
    d.a, d.a_batch_stride, d.a_row_stride, d.a_rows, d.k, d.taps = a.data_ptr(), rows * k, k, rows, k, taps
    if pm is not None:
        d.a_batch_stride, d.a_row_stride = rows * plane_k, plane_k
        d.a_planes, d.a_plane_k, d.a_plane0, d.a_plane_step, d.n_half = n_planes, plane_k, pm.get("plane0", 0), pm.get("step", 1), int(bool(pm.get("n_half")))
    if a_planes is not None:
        d.a_lo = a_planes[1].data_ptr()
    if out_hl16:
        d.c_lo = out[1].data_ptr()
    if addend is not None and addend.dtype == torch.float16:
        d.addend_lo = addend[1].data_ptr()
    d.tap_shift[0], d.tap_shift[1] = shifts
    d.math = _lib.MATH_BY_NAME[math]
    if d.math == _lib.MATH_TF32X3:
        w_hi, w_lo = torch.empty_like(w), torch.empty_like(w)
        check(lib.cum_split_tf32(w.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), w.numel(), _lib.stream_ptr()), "cum_split_tf32")
        d.w, d.w_lo = w_hi.data_ptr(), w_lo.data_ptr()
    elif d.math == _lib.MATH_F16X3:
        import math as _m
        amax = float(w.abs().max())
        e = max(-14, min(14, int(_m.floor(_m.log2(8.0 / amax))))) if amax > 0 else 0
        w_hi, w_lo = torch.empty_like(w, dtype=torch.float16), torch.empty_like(w, dtype=torch.float16)
        check(lib.cum_split_f16(w.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), w.numel(), float(2.0 ** e), _lib.stream_ptr()), "cum_split_f16")
        d.w, d.w_lo, d.acc_scale = w_hi.data_ptr(), w_lo.data_ptr(), float(2.0 ** -e)
    elif d.math == _lib.MATH_BF16X3:
        w_hi, w_lo = torch.empty_like(w, dtype=torch.bfloat16), torch.empty_like(w, dtype=torch.bfloat16)
        check(lib.cum_split_bf16(w.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), w.numel(), _lib.stream_ptr()), "cum_split_bf16")
        d.w, d.w_lo = w_hi.data_ptr(), w_lo.data_ptr()
    else:
        d.w, d.w_lo = w.data_ptr(), 0
    d.ldw, d.bias = k, ptr(bias)
    d.c, d.c_batch_stride, d.c_row_stride, d.m, d.n, d.batch, d.epilogue = out.data_ptr(), m * n_out, n_out, m, n, batch, epilogue
    d.addend, d.add_batch_stride, d.add_row_stride = ptr(addend), m * n_out, n_out
    d.cta_pair = cta_pair
    check(lib.cum_gemm_bias_act_fwd(C.byref(d), _lib.stream_ptr()), "cum_gemm_bias_act_fwd")
    return out
