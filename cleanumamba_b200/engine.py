"""Host-side engine: packs the module's parameters into kernel layouts and runs the forward plan through the C ABI.

Layouts: activations are channels-last fp32 ``(batch, time, C_pad)`` with ``C_pad = ceil8(C)``; weights
are K-contiguous ``(taps, N_pad, K_pad)`` with zero padding, so pad lanes stay exactly 0 through every layer.
  * Conv1d(k=4, s=2)           -> 2-tap GEMM over the ``(L/2, 2 C_pad)`` view of its input, shifts {0, +1}
  * ConvTranspose1d(k=4, s=2)  -> 2-tap GEMM writing the ``(L_in + 1, 2 C_pad)`` view of its output, shifts {0, -1}
  * Conv1d(k=1) + GLU          -> GEMM whose weight rows are interleaved (a_c, b_c) so the gate is a register epilogue
  * U-Net skip add (:315)      -> ``addend`` of the GEMM that produces the decoder-level input
Reference walk: /root/reference/src/network/CleanUMamba.py:252-324.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import functools
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import EPI_GLU, EPI_NONE, EPI_RELU, GemmDesc, ScanDesc, check, ptr

LOG2E = 1.4426950408889634


def p8(n: int) -> int:
    return (n + 7) // 8 * 8


def _pad2(t: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    out = t.new_zeros(rows, cols)
    out[: t.shape[0], : t.shape[1]] = t
    return out


def _pad1(t: torch.Tensor, n: int) -> torch.Tensor:
    out = t.new_zeros(n)
    out[: t.shape[0]] = t
    return out


def _interleave_glu(w: torch.Tensor, b: torch.Tensor, k_pad: int):
    """(2H, K) weight / (2H) bias of a 1x1 conv feeding a GLU -> rows (a_0, b_0, a_1, b_1, ...), padded to ceil8(H)."""
    H = w.shape[0] // 2
    Hp = p8(H)
    wp = w.new_zeros(2 * Hp, k_pad)
    bp = b.new_zeros(2 * Hp)
    wp[0:2 * H:2, : w.shape[1]] = w[:H]
    wp[1:2 * H:2, : w.shape[1]] = w[H:]
    bp[0:2 * H:2] = b[:H]
    bp[1:2 * H:2] = b[H:]
    return wp, bp, Hp


def on_model_device(fn):
    """Run a method with the model's GPU as the current device: the C ABI works on the calling thread's current device and
    ``_lib.stream_ptr()`` is that device's current stream, so a model on cuda:1 works while cuda:0 is current."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self.device_guard():
            return fn(self, *args, **kwargs)
    return wrapper


class Engine:
    skip_zero_lo = True      # skip the a_hi*w_lo pass for weights whose low half is exactly zero (tests may turn it off)
    hl16 = os.environ.get("CUM_HL16", "1") != "0"    # f16x3: store the conv-stack activations as fp16 hi/lo planes (A/B switch)
    HL16_MIN_CONSUMER_K = int(os.environ.get("CUM_HL16_CK", 256))     # thresholds of the per-tensor hl16 rule (see forward())
    HL16_MIN_PRODUCER_K = int(os.environ.get("CUM_HL16_PK", 256))
    fused_ends = os.environ.get("CUM_FUSED_ENDS", "1") != "0"    # f16x3, 64-channel ends: first / last U-Net block as one kernel each
    pack_stream_copies = True  # pack the pitch-32 copies of the conv weights the time-major streaming session uses for irregular widths

    def __init__(self, model):
        self.model = model
        self._key = None
        self.pk: Dict[str, torch.Tensor] = {}
        self.meta: dict = {}
        self._graphs: dict = {}    # (input shape, parameter versions) -> captured CUDA graph of the forward (small problems)
        self.launches = 0          # kernels launched through the C ABI (bench.py's gpu_launches)
        self.prof = None           # list of (kind, start_event, stop_event, flops, bytes) when profiling is on

    def device_guard(self):
        dev = getattr(self, "device", None) or next(self.model.parameters()).device
        if dev.type != "cuda" or dev.index is None or dev.index == torch.cuda.current_device():
            return contextlib.nullcontext()
        return torch.cuda.device(dev)

    def _call(self, kind, fn, *args, launches=1, flops=0, nbytes=0):
        """One C-ABI call on the current stream; optional CUDA-event bracket for per-kernel roofline numbers."""
        self.launches += launches
        if self.prof is None:
            check(fn(*args), fn.__name__)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args), fn.__name__)
        e1.record()
        self.prof.append((kind, e0, e1, flops, nbytes))

    def profile_summary(self):
        """{kind: dict(ms, launches, flops, bytes)} from the recorded events (call after a synchronize)."""
        out = {}
        for kind, e0, e1, fl, nb in self.prof or []:
            d = out.setdefault(kind, dict(ms=0.0, launches=0, flops=0, bytes=0))
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
            d["flops"] += fl
            d["bytes"] += nb
        return out

    # ------------------------------------------------------------------------------------------------ packing
    def _params_key(self):
        return tuple((id(p), p._version, p.data_ptr(), p.device, p.dtype) for p in self.model.parameters())

    pack_generation = 0                # bumped by every repack: sessions holding captured graphs / packed pointers compare it

    def ensure_packed(self):
        key = self._params_key()
        if key != self._key:
            self._graphs.clear()       # captured graphs reference the previous packed weights
            with self.device_guard():
                self._pack()
            self._key = key
            self.pack_generation += 1

    @torch.no_grad()
    def _pack(self):
        m = self.model
        dev = next(m.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("cleanumamba_b200 runs on CUDA (sm_100a) only: move the module with .cuda() "
                               "(there is no CPU fallback)")
        if m.kernel_size != 4 or m.stride != 2:
            raise NotImplementedError("cleanumamba_b200: kernel_size=4, stride=2 only (all shipped configs)")
        if m.channels_input != 1 or m.channels_output != 1:
            raise NotImplementedError("cleanumamba_b200: mono in / mono out only (reference asserts C == 1, :257)")
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32)  # noqa: E731
        D = m.encoder_n_layers
        items: Dict[str, torch.Tensor] = {}
        meta = dict(D=D, enc=[], dec=[], mamba=[])

        c_prev_p = None
        for i, blk in enumerate(m.encoder):
            w0, b0, w1, b1 = f(blk[0].weight), f(blk[0].bias), f(blk[2].weight)[:, :, 0], f(blk[2].bias)
            Hc, Cin, K = w0.shape
            Hc_p = p8(Hc)
            if i == 0:
                items["enc0.w"] = _pad2(w0[:, 0, :].t().contiguous(), K, Hc_p)          # (K, Hc_p) taps-major
            else:
                wt = w0.new_zeros(2, Hc_p, 2 * c_prev_p)
                for s in range(2):
                    for j in range(2):
                        wt[s, :Hc, j * c_prev_p: j * c_prev_p + Cin] = w0[:, :, 2 * s + j]
                items[f"enc{i}.w"] = wt
            items[f"enc{i}.b"] = _pad1(b0, Hc_p)
            wg, bg, Ho_p = _interleave_glu(w1, b1, Hc_p)
            items[f"enc{i}.wg"], items[f"enc{i}.bg"] = wg, bg
            meta["enc"].append(dict(Hc=Hc, Hc_p=Hc_p, Ho=w1.shape[0] // 2, Ho_p=Ho_p, Cin_p=c_prev_p))
            c_prev_p = Ho_p

        w, b = f(m.tsfm_conv1.weight)[:, :, 0], f(m.tsfm_conv1.bias)
        dm = w.shape[0]
        dm_p = p8(dm)
        items["t1.w"], items["t1.b"] = _pad2(w, dm_p, c_prev_p), _pad1(b, dm_p)
        meta.update(dm=dm, dm_p=dm_p, eps=float(m.norm_f.eps))

        for l, blk in enumerate(m.tsfm_Mamba_layers):
            mx = blk.mixer
            di, N = mx.A_log.shape
            R = mx.dt_proj.weight.shape[1]
            W = mx.conv1d.weight.shape[2]
            di_p, N_p, R_p = p8(di), p8(N), p8(R)
            win = f(mx.in_proj.weight)
            wi = win.new_zeros(2 * di_p, dm_p)
            wi[:di, :dm] = win[:di]
            wi[di_p: di_p + di, :dm] = win[di:]
            items[f"m{l}.in"] = wi
            items[f"m{l}.cw"] = _pad2(f(mx.conv1d.weight)[:, 0, :].t().contiguous(), W, di_p)
            items[f"m{l}.cb"] = _pad1(f(mx.conv1d.bias), di_p)
            wx = f(mx.x_proj.weight)
            wxp = wx.new_zeros(R_p + 2 * N_p, di_p)
            wxp[:R, :di] = wx[:R]
            wxp[R_p: R_p + N, :di] = wx[R: R + N]
            wxp[R_p + N_p: R_p + N_p + N, :di] = wx[R + N:]
            items[f"m{l}.xp"] = wxp
            items[f"m{l}.dtw"] = _pad2(f(mx.dt_proj.weight), di_p, R_p)
            items[f"m{l}.dtb"] = _pad1(f(mx.dt_proj.bias), di_p)
            items[f"m{l}.a2"] = _pad2(-torch.exp(f(mx.A_log)) * LOG2E, di_p, N_p)
            items[f"m{l}.D"] = _pad1(f(mx.D), di_p)
            items[f"m{l}.out"] = _pad2(f(mx.out_proj.weight), dm_p, di_p)
            items[f"m{l}.g"], items[f"m{l}.be"] = _pad1(f(blk.norm.weight), dm_p), _pad1(f(blk.norm.bias), dm_p)
            meta["mamba"].append(dict(di=di, di_p=di_p, N=N, N_p=N_p, R=R, R_p=R_p, W=W, eps=float(blk.norm.eps)))
        items["nf.g"], items["nf.be"] = _pad1(f(m.norm_f.weight), dm_p), _pad1(f(m.norm_f.bias), dm_p)

        w, b = f(m.tsfm_conv2.weight)[:, :, 0], f(m.tsfm_conv2.bias)
        c_p = p8(w.shape[0])
        items["t2.w"], items["t2.b"] = _pad2(w, c_p, dm_p), _pad1(b, c_p)
        c_prev_p = c_p
        for j, blk in enumerate(m.decoder):
            w0, b0, wt, bt = f(blk[0].weight)[:, :, 0], f(blk[0].bias), f(blk[2].weight), f(blk[2].bias)
            wg, bg, Hg_p = _interleave_glu(w0, b0, c_prev_p)
            items[f"dec{j}.wg"], items[f"dec{j}.bg"] = wg, bg
            Hg, Co, K = wt.shape
            Co_p = p8(Co)
            if j == D - 1:
                items[f"dec{j}.w"] = _pad2(wt[:, 0, :].t().contiguous(), K, Hg_p)        # (K, Hg_p)
                meta["out_bias"] = float(bt[0].item())
            else:
                wp = wt.new_zeros(2, 2 * Co_p, Hg_p)
                bp = bt.new_zeros(2 * Co_p)
                for s in range(2):
                    for par in range(2):
                        wp[s, par * Co_p: par * Co_p + Co, :Hg] = wt[:, :, 2 * s + par].t()
                for par in range(2):
                    bp[par * Co_p: par * Co_p + Co] = bt
                items[f"dec{j}.w"], items[f"dec{j}.b"] = wp, bp
            meta["dec"].append(dict(Hg=Hg, Hg_p=Hg_p, Co=Co, Co_p=Co_p, Cin_p=c_prev_p))
            c_prev_p = Co_p

        # Time-major streaming (stream_tm.py): a plane-major GEMM operand must be a whole number of 32-element K-blocks wide, so the
        # carried FIFOs use the channel pitch q32(C) and, where that differs from C_pad (pruned checkpoints), the two convolutions get a
        # second packed copy with every K segment padded to the pitch (and the transposed conv's column halves padded to 16)
        q32 = lambda n: (n + 31) // 32 * 32          # noqa: E731
        padded = self.pack_stream_copies
        for i, e in enumerate(meta["enc"]):
            e["Hoq"] = q32(e["Ho_p"])
            if padded and i and q32(e["Cin_p"]) != e["Cin_p"]:
                cp, cq, w = e["Cin_p"], q32(e["Cin_p"]), items[f"enc{i}.w"]
                wq = w.new_zeros(2, w.shape[1], 2 * cq)
                for j in range(2):
                    wq[:, :, j * cq: j * cq + cp] = w[:, :, j * cp: (j + 1) * cp]
                items[f"enc{i}.wq"] = wq
        for j, dd in enumerate(meta["dec"]):
            dd["Hgq"], dd["Coq"] = q32(dd["Hg_p"]), (dd["Co_p"] + 15) // 16 * 16
            if padded and j < D - 1 and (dd["Hgq"] != dd["Hg_p"] or dd["Coq"] != dd["Co_p"]):
                cp, cq, w, b = dd["Co_p"], dd["Coq"], items[f"dec{j}.w"], items[f"dec{j}.b"]
                wq = w.new_zeros(2, 2 * cq, dd["Hgq"])
                for par in range(2):
                    wq[:, par * cq: par * cq + cp, : dd["Hg_p"]] = w[:, par * cp: (par + 1) * cp, :]
                items[f"dec{j}.wq"] = wq
                items[f"dec{j}.bq"] = _pad1(b[:cp], cq)
        items.update(self._extra_items(items))
        # one flat device buffer, every tensor 256-byte aligned
        offs, total = {}, 0
        for k, t in items.items():
            offs[k] = total
            total += (t.numel() + 63) // 64 * 64
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.pk = {}
        for k, t in items.items():
            v = flat[offs[k]: offs[k] + t.numel()].view(t.shape)
            v.copy_(t)
            self.pk[k] = v
        self._flat = flat
        self.meta = meta
        self.device = dev
        mode = getattr(self, "model_math_override", None) or getattr(m, "math_mode", "f16x3")
        self.bf16_io = mode == "bf16"          # reduced-precision variant: bf16 activation storage in the conv stacks
        self.math = _lib.MATH_BY_NAME["f16x3" if self.bf16_io else mode]
        self.pk16 = {k: t.to(torch.bfloat16) for k, t in items.items() if t.dim() >= 2} if self.bf16_io else {}
        self.lib = _lib.init(dev)
        self.pk_hi, self.pk_lo, self.w_scale_inv, self.w_lo_zero = {}, {}, {}, set()
        self.key_math = {}
        if self.math == _lib.MATH_TF32X3:      # hi/lo copies of the whole packed buffer (only GEMM weights use them)
            hi, lo = torch.empty_like(flat), torch.empty_like(flat)
            check(self.lib.cum_split_tf32(flat.data_ptr(), hi.data_ptr(), lo.data_ptr(), flat.numel(),
                                          _lib.stream_ptr()), "cum_split_tf32")
            for k, t in items.items():
                self.pk_hi[k] = hi[offs[k]: offs[k] + t.numel()].view(t.shape)
                self.pk_lo[k] = lo[offs[k]: offs[k] + t.numel()].view(t.shape)
            self._flat_split = (hi, lo)
        elif self.math == _lib.MATH_F16X3:     # fp16 hi/lo copies of 2^k * w, k per weight tensor so that max|2^k w| is in [8, 16)
            hi = torch.empty(flat.numel(), dtype=torch.float16, device=dev)
            lo = torch.empty(flat.numel(), dtype=torch.float16, device=dev)
            gemm_keys = [k for k, t in items.items() if t.dim() == 3 or k.endswith((".wg", ".in", ".xp", ".dtw", ".out", "T")) or k in ("t1.w", "t2.w")]
            amax = torch.stack([items[k].abs().max() for k in gemm_keys]).tolist()          # one host sync at pack time
            for k, a in zip(gemm_keys, amax):
                t = items[k]
                e = int(math.floor(math.log2(8.0 / a))) if a > 0 else 0
                e = max(-14, min(14, e))
                self.w_scale_inv[k] = float(2.0 ** -e)
                check(self.lib.cum_split_f16(flat.data_ptr() + 4 * offs[k], hi.data_ptr() + 2 * offs[k], lo.data_ptr() + 2 * offs[k],
                                             t.numel(), float(2.0 ** e), _lib.stream_ptr()), "cum_split_f16")
                self.pk_hi[k] = hi[offs[k]: offs[k] + t.numel()].view(t.shape)
                self.pk_lo[k] = lo[offs[k]: offs[k] + t.numel()].view(t.shape)
            self._flat_split = (hi, lo)
        elif self.math == _lib.MATH_BF16X3:    # bf16 hi/lo copies (same element offsets, 2 bytes per element)
            hi = torch.empty(flat.numel(), dtype=torch.bfloat16, device=dev)
            lo = torch.empty(flat.numel(), dtype=torch.bfloat16, device=dev)
            check(self.lib.cum_split_bf16(flat.data_ptr(), hi.data_ptr(), lo.data_ptr(), flat.numel(),
                                          _lib.stream_ptr()), "cum_split_bf16")
            for k, t in items.items():
                self.pk_hi[k] = hi[offs[k]: offs[k] + t.numel()].view(t.shape)
                self.pk_lo[k] = lo[offs[k]: offs[k] + t.numel()].view(t.shape)
            self._flat_split = (hi, lo)
        if self.pk_lo:
            # weights that are exactly representable by their high half (e.g. a checkpoint shipped in fp16, loaded with
            # .float(), under f16x3): the third MMA pass multiplies by zeros and is skipped (identical result)
            keys = [k for k in self.pk_lo if self.pk_lo[k].numel() > 0]
            nz = torch.stack([self.pk_lo[k].float().abs().max() for k in keys]).tolist()     # one host sync at pack time
            self.w_lo_zero = {k for k, v in zip(keys, nz) if v == 0.0}
        self._post_pack(items, offs, total)

    def _extra_items(self, items):
        """Hook: additional packed tensors (TrainEngine adds the transposed weights used by the data gradients)."""
        return {}

    def _post_pack(self, items, offs, total):
        """Hook: called once the packed buffer exists (TrainEngine allocates the gradient buffer here)."""

    # ------------------------------------------------------------------------------------------------ op wrappers
    def gemm(self, a, a_off, a_bs, a_rs, a_rows, k, w, bias, c, c_off, c_bs, c_rs, m, n, batch, epi,
             taps=1, shifts=(0, 0), addend=None, add_bs=0, add_rs=0, math=None, a_scale=None, aux=None, aux_bs=0, aux_rs=0, addend_mask=False,
             planes=None, add_off=0, small=0):
        """``planes`` = (a_planes, a_plane_k, a_plane0, a_plane_step, n_half): plane-major ``a`` (cum_gemm_desc.a_planes; the time-major
        streaming session).  ``add_off``: element offset into ``addend``.  ``small``: cum_gemm_desc.small_m_path (the streaming sessions
        decide per level from columns x streams, so that both buffer layouts run the same kernel for the same level)."""
        d = GemmDesc()
        d.small_m_path = small
        if planes is not None:
            d.a_planes, d.a_plane_k, d.a_plane0, d.a_plane_step, d.n_half = planes
        d.a, d.a_batch_stride, d.a_row_stride, d.a_rows, d.k = a.data_ptr() + a.element_size() * a_off, a_bs, a_rs, a_rows, k
        # "hl16" tensors (dtype float16, leading dimension 2 = [hi plane, lo plane], see cum_gemm_desc.a_lo): pre-split activations
        if a.dtype == torch.float16:
            d.a_lo = a[1].data_ptr() + 2 * a_off
        if c.dtype == torch.float16:
            d.c_lo = c[1].data_ptr() + 2 * c_off
        if addend is not None and addend.dtype == torch.float16:
            d.addend_lo = addend[1].data_ptr()
        d.taps = taps
        d.tap_shift[0], d.tap_shift[1] = shifts
        wt = self.pk[w]
        d.math = (self.key_math.get(w, self.math) if math is None else math)      # per-weight override (TrainEngine: dgrad weights)
        d.out_bf16 = 1 if c.dtype == torch.bfloat16 else 0
        if a.dtype == torch.bfloat16:          # reduced-precision variant: bf16 activations x bf16 weights, one MMA pass
            d.math = _lib.MATH_BF16
            d.w, d.w_lo = self.pk16[w].data_ptr(), 0
        elif d.math in (_lib.MATH_TF32X3, _lib.MATH_BF16X3, _lib.MATH_F16X3):
            d.w, d.w_lo = self.pk_hi[w].data_ptr(), self.pk_lo[w].data_ptr()
            d.acc_scale = self.w_scale_inv.get(w, 1.0)
            d.w_lo_is_zero = 1 if (self.skip_zero_lo and w in self.w_lo_zero) else 0
        else:
            d.w, d.w_lo = wt.data_ptr(), 0
        d.ldw, d.bias = wt.shape[-1], ptr(bias)
        d.c, d.c_batch_stride, d.c_row_stride, d.m, d.n, d.batch = c.data_ptr() + c.element_size() * c_off, c_bs, c_rs, m, n, batch
        d.epilogue = epi
        d.addend, d.add_batch_stride, d.add_row_stride = ptr(addend), add_bs, add_rs
        if addend is not None and add_off:
            d.addend = addend.data_ptr() + addend.element_size() * add_off
        d.addend_is_mask = 1 if (addend_mask and addend is not None) else 0
        if a_scale is not None and d.math == _lib.MATH_F16X3 and a.dtype == torch.float32:
            d.a_scale_dev = a_scale.data_ptr()          # device-side power-of-two scale of a gradient operand (TrainEngine)
        if aux is not None:         # training: second fp32 output (pre-gate / pre-addend value), see cum_gemm_desc.aux
            d.aux, d.aux_batch_stride, d.aux_row_stride = aux.data_ptr(), aux_bs, aux_rs
        kind = "gemm_tap2" if taps == 2 else "gemm"
        self._call(kind, self.lib.cum_gemm_bias_act_fwd, C.byref(d), _lib.stream_ptr(),
                   flops=2 * batch * m * n * k * taps)

    def dense(self, a, rows, k, w, bias, n, epi=EPI_NONE, a_rs=None, a_off=0, addend=None, out=None, out_dtype=torch.float32, aux=None, out_off=0,
              out_rs=None, add_rs=None, small=0):
        """Flat (rows, k) x W^T -> (rows, n or n/2): 1x1 convs and Linear layers.  ``aux`` (rows, n): the pre-activation (training).
        ``a_rs`` / ``out_rs`` / ``add_rs``: row pitches of the operands when they are wider than their logical width."""
        n_out = n // 2 if epi >= 8 else n
        c = out if out is not None else self.act_buffer(rows, n_out, out_dtype, a.device)
        self.gemm(a, a_off, 0, k if a_rs is None else a_rs, rows, k, w, bias, c, out_off, 0, n_out if out_rs is None else out_rs, rows, n, 1, epi,
                  addend=addend, add_bs=0, add_rs=n_out if add_rs is None else add_rs, aux=aux, aux_rs=n, small=small)
        return c

    @staticmethod
    def act_buffer(rows, c, dtype, device):
        """Activation storage: fp32 / bf16 (rows, c), or for float16 the two-plane hl16 format (2, rows, c)."""
        if dtype == torch.float16:
            return torch.empty(2, rows, c, dtype=dtype, device=device)
        return torch.empty(rows, c, dtype=dtype, device=device)

    def ln(self, h, res_in, res_out, normed, g, be, eps, rows, c, c_p):
        self._call("ln_residual", self.lib.cum_ln_residual_fwd, ptr(h), ptr(res_in), ptr(res_out), ptr(normed), ptr(g),
                   ptr(be), eps, rows, c, c_p, _lib.stream_ptr(),
                   nbytes=4 * rows * c * (2 + (res_in is not None) + (res_out is not None)))

    def fill_scan(self, s, u, dt, xz, xdbl, y, l, mm, B, T, h0=None, h_out=None, h_ckpt=None, tm=False):
        """Fill a cum_scan_desc for Mamba layer ``l`` operating on the engine's channels-last buffers.  ``tm``: the rows are
        time-major (token t of stream b is row t * B + b: the streaming session for many streams) instead of (b, t)."""
        di_p, N_p, R_p = mm["di_p"], mm["N_p"], mm["R_p"]
        ld = R_p + 2 * N_p

        def strides(width):            # (batch stride, row stride) of a (rows, width) array
            return (width, B * width) if tm else (T * width, width)
        s.u, (s.u_bs, s.u_rs) = u.data_ptr(), strides(di_p)
        s.delta, (s.dl_bs, s.dl_rs) = dt.data_ptr(), strides(di_p)
        s.z, (s.z_bs, s.z_rs) = xz.data_ptr() + 4 * di_p, strides(2 * di_p)
        s.Bm, (s.B_bs, s.B_rs) = xdbl.data_ptr() + 4 * R_p, strides(ld)
        s.Cm, (s.C_bs, s.C_rs) = xdbl.data_ptr() + 4 * (R_p + N_p), strides(ld)
        s.y, (s.y_bs, s.y_rs) = ptr(y), strides(di_p)
        s.a2, s.Dskip, s.delta_bias = self.pk[f"m{l}.a2"].data_ptr(), self.pk[f"m{l}.D"].data_ptr(), self.pk[f"m{l}.dtb"].data_ptr()
        s.h0, s.h_out, s.h_ckpt = ptr(h0), ptr(h_out), ptr(h_ckpt)
        s.state_f16 = 1 if (h0 is not None and h0.dtype == torch.float16) else 0
        s.batch, s.len, s.d, s.n_state, s.delta_softplus = B, T, di_p, N_p, 1

    def scan(self, u, dt, xz, xdbl, y, l, mm, B, T, h0=None, h_out=None, h_ckpt=None, tm=False):
        s = ScanDesc()
        self.fill_scan(s, u, dt, xz, xdbl, y, l, mm, B, T, h0, h_out, h_ckpt, tm=tm)
        ws = None
        if h_ckpt is None and T >= 256:        # small batches of long clips: segment-parallel scan (scratch from the caching allocator)
            from .ops import scan_workspace
            ws = scan_workspace(self.lib, s, u.device)
        # algorithmic bytes: read u, delta, z + B, C, write y -- real (unpadded) widths
        self._call("selective_scan", self.lib.cum_selective_scan_fwd, C.byref(s), _lib.stream_ptr(),
                   nbytes=4 * B * T * (4 * mm["di"] + 2 * mm["N"]), flops=B * T * mm["di"] * mm["N"], launches=1 if ws is None else 3)
        del ws

    def mamba_layers(self, h, B, T, states=None, tm=False, small=0):
        """h: (B*T, dm_p) output of tsfm_conv1 -> normed (B*T, dm_p) after norm_f.  ``states``: optional list of
        (conv_state (B, W-1, di_p), ssm_state (B, di_p, N_p)) carried in place (streaming; an fp16 ssm_state selects the
        reduced-precision state variant).  ``tm``: rows are time-major (t * B + b) instead of (b * T + t)."""
        pk, meta = self.pk, self.meta
        dm, dm_p = meta["dm"], meta["dm_p"]
        rows = B * T
        dev = h.device
        res = None
        hn = torch.empty(rows, dm_p, dtype=torch.float32, device=dev)
        for l, mm in enumerate(meta["mamba"]):
            di_p, N_p, R_p = mm["di_p"], mm["N_p"], mm["R_p"]
            res_out = res if res is not None else torch.empty(rows, dm_p, dtype=torch.float32, device=dev)
            self.ln(h, res, res_out, hn, pk[f"m{l}.g"], pk[f"m{l}.be"], mm["eps"], rows, dm, dm_p)
            res = res_out
            xz = self.dense(hn, rows, dm_p, f"m{l}.in", None, 2 * di_p, small=small)
            xc = torch.empty(rows, di_p, dtype=torch.float32, device=dev)
            cs = states[l][0] if states is not None else None
            n_l = 1 + (cs is not None and T > 16)      # up to 16 tokens the kernel writes the new conv state itself
            if tm:
                self._call("dwconv_silu", self.lib.cum_dwconv_silu_strided_fwd, xz.data_ptr(), 2 * di_p, B * 2 * di_p,
                           pk[f"m{l}.cw"].data_ptr(), pk[f"m{l}.cb"].data_ptr(), xc.data_ptr(), di_p, B * di_p, ptr(cs), ptr(cs), B, T, di_p,
                           mm["W"], _lib.stream_ptr(), launches=n_l, nbytes=8 * rows * mm["di"])
            else:
                self._call("dwconv_silu", self.lib.cum_dwconv_silu_fwd, xz.data_ptr(), T * 2 * di_p, 2 * di_p,
                           pk[f"m{l}.cw"].data_ptr(), pk[f"m{l}.cb"].data_ptr(), xc.data_ptr(), ptr(cs), ptr(cs), B, T, di_p,
                           mm["W"], _lib.stream_ptr(), launches=n_l, nbytes=8 * rows * mm["di"])
            xdbl = self.dense(xc, rows, di_p, f"m{l}.xp", None, R_p + 2 * N_p, small=small)
            dt = self.dense(xdbl, rows, R_p, f"m{l}.dtw", None, di_p, a_rs=R_p + 2 * N_p, small=small)
            y = torch.empty(rows, di_p, dtype=torch.float32, device=dev)
            hs = states[l][1] if states is not None else None
            self.scan(xc, dt, xz, xdbl, y, l, mm, B, T, h0=hs, h_out=hs, tm=tm)
            h = self.dense(y, rows, di_p, f"m{l}.out", None, dm_p, small=small)
        self.ln(h, res, None, hn, pk["nf.g"], pk["nf.be"], meta["eps"], rows, dm, dm_p)
        return hn

    # ------------------------------------------------------------------------------------------------ forward
    # Small problems (the pruned 200K-2M checkpoints at batch 1-8: ~285 launches in ~1.4 ms): after GRAPH_AFTER eager calls with
    # the same input shape and the same parameter versions the whole forward is captured once as a CUDA graph and replayed
    # (static input / activation buffers; bit-identical).  Measured gain is modest (E8-500K, 1 x 10 s: 1.42 -> 1.32 ms): the
    # time is the start-up cost of 285 persistent kernels, not host launch latency -- the real fix for these tiny irregular
    # layers is a fused multi-layer kernel.  Larger problems are GPU-bound and stay eager.
    GRAPH_MAX_SAMPLES = int(os.environ.get("CUM_GRAPH_MAX_SAMPLES", 8 * 160000))     # batch x samples up to which graphs are used
    GRAPH_AFTER = 2
    GRAPH_CACHE = 4

    @torch.no_grad()
    @on_model_device
    def forward(self, noisy: torch.Tensor, return_skip_connections: bool = False):
        self.ensure_packed()
        if (return_skip_connections or self.prof is not None or noisy.dim() != 3 or noisy.numel() > self.GRAPH_MAX_SAMPLES
                or noisy.device != self.device or torch.cuda.is_current_stream_capturing()):
            return self._forward_eager(noisy, return_skip_connections)
        key = (tuple(noisy.shape), self._key, self.model.normalize_input, self.model.glu_activation)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = dict(seen=0, graph=None)
        if ent["graph"] is None:
            ent["seen"] += 1
            if ent["seen"] <= self.GRAPH_AFTER:
                return self._forward_eager(noisy, False)
            x_static = torch.empty(noisy.shape, dtype=torch.float32, device=self.device)
            x_static.copy_(noisy)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward_eager(x_static, False)
            ent.update(graph=g, x=x_static, out=out)
        ent["x"].copy_(noisy)
        ent["graph"].replay()
        if self.model.normalize_input:
            noisy.copy_(ent["x"])            # the reference divides the caller's tensor in place (:262)
        return ent["out"].clone()

    @torch.no_grad()
    def _forward_eager(self, noisy: torch.Tensor, return_skip_connections: bool = False):
        m = self.model
        self.ensure_packed()
        pk, meta, lib = self.pk, self.meta, self.lib
        if noisy.device != self.device:
            raise RuntimeError(f"input is on {noisy.device}, model on {self.device} (no CPU fallback)")
        B, _, L = noisy.shape
        D = meta["D"]
        st = _lib.stream_ptr
        act = EPI_GLU[m.glu_activation]

        x = noisy
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = noisy.to(torch.float32).contiguous()
        std = None
        if m.normalize_input:
            std = torch.empty(B, dtype=torch.float32, device=x.device)
            self._call("wave_normalize", lib.cum_wave_normalize_fwd, x.data_ptr(), std.data_ptr(), B, L, st(),
                       nbytes=8 * B * L)
            if x is not noisy:
                noisy.copy_(x)      # the reference divides the caller's tensor in place (:262)
        Ls = [m.valid_length(L)]
        for _ in range(D):
            Ls.append((Ls[-1] - 4) // 2 + 1)

        # storage of the encoder / decoder activations: bf16 (reduced-precision variant); fp16 hi/lo planes under f16x3 (each
        # GEMM epilogue splits its output once, the consumer needs no operand splitter -- same products as fp32 storage); else fp32
        hl16 = (not self.bf16_io) and self.math == _lib.MATH_F16X3 and self.hl16
        adt = torch.bfloat16 if self.bf16_io else torch.float32

        def fmt(consumer_k, producer_k):
            """Per tensor: hl16 pays when the consuming GEMM is compute-bound (it loses its splitter: -6..13 % at K >= 512) and
            the producing GEMM is not epilogue-bound (the split store costs it +3..5 % at K >= 256, +25 % on the 128-wide layers)"""
            return torch.float16 if (hl16 and consumer_k >= self.HL16_MIN_CONSUMER_K and producer_k >= self.HL16_MIN_PRODUCER_K) else adt
        skips: List[torch.Tensor] = []
        prev = None
        fuse = (self.fused_ends and self.math == _lib.MATH_F16X3 and not self.bf16_io and m.glu_activation == "Sigmoid" and D > 1)
        e0, dl = meta["enc"][0], meta["dec"][D - 1]
        fuse_e0 = fuse and e0["Hc_p"] <= 64 and e0["Ho_p"] <= 64          # narrower (pruned) layers run zero-padded in the 64 x 128 tile
        fuse_dl = fuse and dl["Cin_p"] <= 64 and dl["Hg_p"] <= 64
        for i, e in enumerate(meta["enc"]):
            rows = B * Ls[i + 1]
            if i == 0 and fuse_e0:
                # whole first block in one kernel: the 64-channel conv output stays on the SM (fused_ends.cu)
                prev = torch.empty(rows, e["Ho_p"], dtype=torch.float32, device=x.device)
                d0 = _lib.Enc0BlockDesc()
                d0.x, d0.x_stride, d0.batch, d0.length = x.data_ptr(), L, B, L
                d0.conv_w, d0.conv_b = pk["enc0.w"].data_ptr(), pk["enc0.b"].data_ptr()
                d0.glu_w_hi, d0.glu_w_lo, d0.glu_b = self.pk_hi["enc0.wg"].data_ptr(), self.pk_lo["enc0.wg"].data_ptr(), pk["enc0.bg"].data_ptr()
                d0.acc_scale = self.w_scale_inv["enc0.wg"]
                d0.w_lo_is_zero = 1 if (self.skip_zero_lo and "enc0.wg" in self.w_lo_zero) else 0
                d0.out, d0.rows_out, d0.channels, d0.channels_out = prev.data_ptr(), Ls[1], e["Hc_p"], e["Ho_p"]
                self._call("enc0_block", lib.cum_enc0_block_fwd, C.byref(d0), st(), flops=2 * rows * 2 * e["Ho_p"] * e["Hc_p"],
                           nbytes=4 * B * (L + Ls[1] * e["Ho"]))
                skips.append(prev)
                continue
            y = self.act_buffer(rows, e["Hc_p"], fmt(e["Hc_p"], 2 * e["Cin_p"] if i else 0), x.device)
            if i == 0 and y.dtype == torch.float16:
                self._call("conv_in", lib.cum_conv_in_hl16_fwd, x.data_ptr(), L, B, L, pk["enc0.w"].data_ptr(),
                           pk["enc0.b"].data_ptr(), y[0].data_ptr(), y[1].data_ptr(), Ls[1], e["Hc_p"], 4, 2, st(),
                           nbytes=4 * B * (L + Ls[1] * e["Hc"]))
            elif i == 0 and self.bf16_io:
                self._call("conv_in", lib.cum_conv_in_bf16_fwd, x.data_ptr(), L, B, L, pk["enc0.w"].data_ptr(),
                           pk["enc0.b"].data_ptr(), y.data_ptr(), Ls[1], e["Hc_p"], 4, 2, st(),
                           nbytes=B * (4 * L + 2 * Ls[1] * e["Hc"]))
            elif i == 0:
                self._call("conv_in", lib.cum_conv_in_fwd, x.data_ptr(), L, B, L, pk["enc0.w"].data_ptr(),
                           pk["enc0.b"].data_ptr(), y.data_ptr(), Ls[1], e["Hc_p"], 4, 2, 0, 0, 0, st(),
                           nbytes=4 * B * (L + Ls[1] * e["Hc"]))
            else:
                cp = e["Cin_p"]
                self.gemm(prev, 0, Ls[i] * cp, 2 * cp, Ls[i] // 2, 2 * cp, f"enc{i}.w", pk[f"enc{i}.b"],
                          y, 0, Ls[i + 1] * e["Hc_p"], e["Hc_p"], Ls[i + 1], e["Hc_p"], B, EPI_RELU, taps=2,
                          shifts=(0, 1))
            prev = self.dense(y, rows, e["Hc_p"], f"enc{i}.wg", pk[f"enc{i}.bg"], 2 * e["Ho_p"], epi=act,
                              out_dtype=fmt(2 * e["Ho_p"] if i < D - 1 else e["Ho_p"], e["Hc_p"]))
            skips.append(prev)

        T = Ls[D]
        rows = B * T
        cb_p = meta["enc"][-1]["Ho_p"]
        h = self.dense(prev, rows, cb_p, "t1.w", pk["t1.b"], meta["dm_p"])
        hn = self.mamba_layers(h, B, T)
        xcur = self.dense(hn, rows, meta["dm_p"], "t2.w", pk["t2.b"], cb_p, addend=skips[D - 1], out_dtype=fmt(cb_p, meta["dm_p"]))

        Tj = T
        out = None
        for j, d in enumerate(meta["dec"]):
            if j == D - 1 and fuse_dl and xcur.dtype == torch.float32:
                # whole last block in one kernel: GLU GEMM + transposed conv to the waveform + crop + * std (fused_ends.cu)
                length = L if m.normalize_input else Ls[0]
                out = torch.empty(B, 1, length, dtype=torch.float32, device=x.device)
                dd = _lib.DecLastBlockDesc()
                dd.a, dd.batch, dd.rows_in = xcur.data_ptr(), B, Tj
                dd.glu_w_hi, dd.glu_w_lo, dd.glu_b = self.pk_hi[f"dec{j}.wg"].data_ptr(), self.pk_lo[f"dec{j}.wg"].data_ptr(), pk[f"dec{j}.bg"].data_ptr()
                dd.acc_scale = self.w_scale_inv[f"dec{j}.wg"]
                dd.w_lo_is_zero = 1 if (self.skip_zero_lo and f"dec{j}.wg" in self.w_lo_zero) else 0
                dd.convt_w, dd.convt_bias, dd.scale = pk[f"dec{j}.w"].data_ptr(), meta["out_bias"], ptr(std)
                dd.out, dd.out_stride, dd.out_length, dd.channels, dd.channels_gated = out.data_ptr(), length, length, d["Cin_p"], d["Hg_p"]
                self._call("dec_last_block", lib.cum_dec_last_block_fwd, C.byref(dd), st(), flops=2 * B * Tj * 2 * d["Hg_p"] * d["Cin_p"],
                           nbytes=4 * B * (Tj * d["Hg"] + length))
                break
            # (the waveform-end kernel reads plain fp32 / bf16 rows: the last GLU output is never written as hi/lo planes)
            g = self.dense(xcur, B * Tj, d["Cin_p"], f"dec{j}.wg", pk[f"dec{j}.bg"], 2 * d["Hg_p"], epi=act,
                           out_dtype=fmt(2 * d["Hg_p"], d["Cin_p"]) if j < D - 1 else adt)
            if j < D - 1:
                co = d["Co_p"]
                To = 2 * Tj + 2
                nxt = self.act_buffer(B * To, co, fmt(co, 2 * d["Hg_p"]), x.device)
                skip = skips[D - 2 - j]
                self.gemm(g, 0, Tj * d["Hg_p"], d["Hg_p"], Tj, d["Hg_p"], f"dec{j}.w", pk[f"dec{j}.b"],
                          nxt, 0, To * co, 2 * co, Tj + 1, 2 * co, B, EPI_RELU, taps=2, shifts=(0, -1),
                          addend=skip, add_bs=To * co, add_rs=2 * co)
                xcur, Tj = nxt, To
            else:
                length = L if m.normalize_input else Ls[0]
                out = torch.empty(B, 1, length, dtype=torch.float32, device=x.device)
                self._call("convt_out", lib.cum_convt_out_bf16_fwd if self.bf16_io else lib.cum_convt_out_fwd,
                           g.data_ptr(), B, Tj, d["Hg_p"], pk[f"dec{j}.w"].data_ptr(),
                           meta["out_bias"], ptr(std), length, out.data_ptr(), length, 0, length, 4, 2, st(),
                           nbytes=4 * B * (Tj * d["Hg"] + length))
        if not return_skip_connections:
            return out
        ncl = []
        for i in reversed(range(D)):      # the reference returns the skips deepest-first (:275) in (B, C, L)
            e = meta["enc"][i]
            sk = (skips[i][0].float() + skips[i][1].float()) if skips[i].dtype == torch.float16 else skips[i]
            ncl.append(sk.view(B, Ls[i + 1], e["Ho_p"])[:, :, : e["Ho"]].permute(0, 2, 1).float())
        ncl.append(hn.view(B, T, meta["dm_p"])[:, :, : meta["dm"]].permute(0, 2, 1))
        return out, ncl
