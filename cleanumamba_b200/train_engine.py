"""Training-mode forward (activations saved) and backward of the CleanUMamba path through the C ABI.

Reference: autograd through /root/reference/src/network/CleanUMamba.py:252-324 as driven by the training step
(src/training/train.py:278-285).  Differences from the inference plan (engine.py): the GLU gate and the U-Net skip add
run as their own kernels so each pre-activation is stored exactly once; the scan additionally writes its chunk-start
states (h_ckpt).  Every data gradient of a dense layer is the forward tap-GEMM with transposed packed weights and
negated tap shifts; weight gradients use cum_gemm_wgrad (fp32).  Gradients are accumulated in a flat buffer with the
packed-weight layout and unpacked into parameter shapes at the end (``unpack_grads``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

import torch

from . import _lib
from ._lib import EPI_GLU, EPI_NONE, EPI_RELU, ScanBwdDesc, ScanDesc, WgradDesc, check, ptr
from .engine import Engine, on_model_device


class TrainEngine(Engine):
    """Engine with transposed weight copies (for dgrad) and a gradient buffer.

    Under "f16x3" every GEMM of a training step uses the fp16 split (2x the TF32 tensor rate).  Back-propagated gradients
    routinely fall below fp16's 6e-5 normal range, so each gradient tensor that feeds a GEMM is scaled by a power of two
    computed on the device from its max (``grad_scale``): the data-gradient GEMM's operand splitter multiplies by it and the
    epilogue divides it out, the weight-gradient GEMM (MN-major operands, no transposes) shares it.  ``CUM_TRAIN_F16_BWD=0``
    restores the round-1 arithmetic (transposed weights as TF32 halves, "tf32x3" data gradients).  "bf16" (inference-only
    storage variant) maps to "tf32x3" entirely."""

    tc_wgrad = True      # tensor-core weight gradients (tests may switch to the fp32 CUDA-core kernel)
    # f16x3 models: the data-gradient GEMMs run in f16x3 as well (twice the TF32 tensor rate).  Back-propagated gradients (~1e-6)
    # underflow fp16, so every gradient tensor that feeds a GEMM gets a power-of-two scale computed ON THE DEVICE from its max
    # (cum_grad_scale_fwd): the operand splitter multiplies by it, the epilogue divides it out, the weight-gradient GEMM shares it.
    f16_backward = os.environ.get("CUM_TRAIN_F16_BWD", "1") != "0"
    # tensor-core modes: the GLU gate and the decoder's skip add run in the GEMM epilogue, which ALSO stores what the backward needs
    # (the pre-activation / the pre-add value: cum_gemm_desc.aux) -- no stand-alone glu_fwd / add kernels
    fused_forward = os.environ.get("CUM_TRAIN_FUSED_FWD", "1") != "0"
    f16_forward = os.environ.get("CUM_TRAIN_F16_FWD", "1") != "0"     # f16x3 models: forward GEMMs in f16x3 (A/B switch)
    # ReLU backward as a mask in the epilogue of the data-gradient GEMM that produces its input gradient (cum_gemm_desc.addend_is_mask).
    # Measured (same box, 16 x 10 s): 55.3 ms with it, 54.8 ms without -- the mask read + the separate column-sum pass cost more than
    # the stand-alone relu_bwd kernel saves; off by default.
    fused_relu_bwd = os.environ.get("CUM_TRAIN_FUSED_RELU_BWD", "0") != "0"

    def _pack(self):
        if getattr(self.model, "math_mode", None) == "bf16" or not self.f16_forward and getattr(self.model, "math_mode", None) == "f16x3":
            self.model_math_override = "tf32x3"
        super()._pack()

    pack_stream_copies = False     # training never streams

    def _extra_items(self, items: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        extra = {}
        for k, t in items.items():
            if t.dim() == 3 and not k.endswith(".w0"):        # (taps, N, K) tap-GEMM weights
                extra[k + "T"] = t.transpose(1, 2).contiguous()
            elif t.dim() == 2 and (k.endswith((".wg", ".in", ".xp", ".dtw", ".out")) or k in ("t1.w", "t2.w")):
                extra[k + "T"] = t.t().contiguous()
        return extra

    def _post_pack(self, items, offs, total):
        self._scale_pool = torch.zeros(128, 4, dtype=torch.float32, device=self.device)
        self._scale_next = 0
        if self.math == _lib.MATH_F16X3 and not self.f16_backward:
            # transposed (dgrad) weights: TF32 hi/lo halves instead of the fp16 split the base class prepared for every GEMM weight
            tkeys = [k for k in items if k.endswith("T")]
            lo0 = min(offs[k] for k in tkeys)
            hi0 = max(offs[k] + (items[k].numel() + 63) // 64 * 64 for k in tkeys)
            thi = torch.empty(hi0 - lo0, dtype=torch.float32, device=self.device)
            tlo = torch.empty_like(thi)
            check(self.lib.cum_split_tf32(self._flat.data_ptr() + 4 * lo0, thi.data_ptr(), tlo.data_ptr(), hi0 - lo0, _lib.stream_ptr()),
                  "cum_split_tf32")
            for k in tkeys:
                o = offs[k] - lo0
                self.pk_hi[k] = thi[o: o + items[k].numel()].view(items[k].shape)
                self.pk_lo[k] = tlo[o: o + items[k].numel()].view(items[k].shape)
                self.key_math[k] = _lib.MATH_TF32X3
                self.w_scale_inv.pop(k, None)
                self.w_lo_zero.discard(k)
            self._t_split = (thi, tlo)
        # gradient buffer: same offsets as the packed weights for the base items (they precede the transposed copies),
        # followed by a slot for the scalar bias of the last transposed conv
        base = [k for k in items if not k.endswith("T")]
        base_total = max(offs[k] + (items[k].numel() + 63) // 64 * 64 for k in base)
        self.gflat = torch.zeros(base_total + 64, dtype=torch.float32, device=self.device)
        self.gk = {k: self.gflat[offs[k]: offs[k] + items[k].numel()].view(items[k].shape) for k in base}
        self.gk["out_bias"] = self.gflat[base_total: base_total + 4]
        # all-reduce buckets in the order the backward completes them: decoder (+ out bias), bottleneck, the DEEP encoder levels (almost
        # all of the encoder's bytes: their reduction hides behind the backward of the outer levels, whose activations are the largest
        # of the model), and last the few outer levels (<= 2 MB: all that is left exposed)
        D = self.meta["D"]
        k_split = 1
        while k_split < D - 1 and 4 * offs[f"enc{k_split + 1}.w"] <= (2 << 20):
            k_split += 1
        self.enc_split_level = k_split if D >= 2 else -1
        cut = offs[f"enc{k_split}.w"] if D >= 2 else offs["t1.w"]
        self.buckets = dict(decoder=(offs["dec0.wg"], base_total + 64), bottleneck=(offs["t1.w"], offs["dec0.wg"]),
                            encoder_deep=(cut, offs["t1.w"]), encoder_outer=(0, cut))

    # ---------------------------------------------------------------------------------------------- helpers
    def new(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)

    def scale_slot(self):
        """A free {s, 1/s, -, -} slot for a kernel that computes the scale of the gradient it produces (glu_bwd / relu_bwd)."""
        if self.math != _lib.MATH_F16X3 or not self.f16_backward:
            return None
        sc = self._scale_pool[self._scale_next % self._scale_pool.shape[0]]
        self._scale_next += 1
        return sc

    def grad_scale(self, t, rows, cols):
        """{s, 1/s} (device, 4 floats) of the contiguous (rows, cols) gradient tensor ``t``: s = the power of two that lifts max|t| to
        [2^14, 2^15).  None outside f16x3 (the other modes keep fp32 range)."""
        if self.math != _lib.MATH_F16X3 or not self.f16_backward:
            return None
        sc = self._scale_pool[self._scale_next % self._scale_pool.shape[0]]
        self._scale_next += 1
        self._call("grad_scale", self.lib.cum_grad_scale_fwd, t.data_ptr(), 0, cols, 1, rows, cols, sc.data_ptr(), _lib.stream_ptr(),
                   launches=2, nbytes=4 * rows * cols)
        return sc

    def wgrad(self, dz, dz_bs, dz_rs, a, a_off, a_bs, a_rs, a_rows, key, m, n, k, batch, taps=1, shifts=(0, 0), scale=None):
        d = WgradDesc()
        d.dz, d.dz_batch_stride, d.dz_row_stride = dz.data_ptr(), dz_bs, dz_rs
        d.a, d.a_batch_stride, d.a_row_stride, d.a_rows = a.data_ptr() + 4 * a_off, a_bs, a_rs, a_rows
        g = self.gk[key]
        d.dw, d.ldw = g.data_ptr(), g.shape[-1]
        d.m, d.n, d.k, d.taps, d.batch = m, n, k, taps, batch
        d.tap_shift[0], d.tap_shift[1] = shifts
        d.math = _lib.MATH_FP32 if (self.math == _lib.MATH_FP32 or not self.tc_wgrad) else _lib.MATH_TF32X3
        if scale is not None:
            d.dz_scale_dev = scale.data_ptr()
        ws = None
        if d.math != _lib.MATH_FP32:
            nbytes = self.lib.cum_gemm_wgrad_workspace_bytes(C.byref(d))
            ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            d.workspace = ws.data_ptr()
        self._call("wgrad", self.lib.cum_gemm_wgrad, C.byref(d), _lib.stream_ptr(), flops=2 * batch * m * n * k * taps,
                   launches=1 if ws is None else 2 + taps + (0 if scale is not None else 3))

    def dense_T(self, dz, rows, n_fwd, key, k_fwd, addend=None, out=None, c_rs=None, scale=None, relu_mask=None, unmasked=None):
        """Data gradient of a flat dense layer: (rows, n_fwd) x W (n_fwd, k_fwd) -> (rows, k_fwd).  ``relu_mask`` (rows, k_fwd): the
        forward's ReLU output -- the gradient is zeroed where it is not positive (ReLU backward in the epilogue), ``unmasked``
        receives the gradient before the mask."""
        c = out if out is not None else self.new(rows, k_fwd)
        crs = k_fwd if c_rs is None else c_rs
        if relu_mask is not None:
            self.gemm(dz, 0, 0, n_fwd, rows, n_fwd, key + "T", None, c, 0, 0, crs, rows, k_fwd, 1, EPI_NONE,
                      addend=relu_mask, add_bs=0, add_rs=crs, addend_mask=True, a_scale=scale, aux=unmasked, aux_rs=crs)
        else:
            self.gemm(dz, 0, 0, n_fwd, rows, n_fwd, key + "T", None, c, 0, 0, crs, rows, k_fwd, 1, EPI_NONE,
                      addend=addend, add_bs=0, add_rs=crs, a_scale=scale)
        return c

    def colsum_scale(self, t, bias_key, rows, cols):
        """Bias gradient (column sums) of the gradient tensor ``t`` and, in the same pass, its device scale for the f16x3 GEMMs."""
        sc = self.scale_slot()
        self._call("colsum", self.lib.cum_colsum, t.data_ptr(), self.gk[bias_key].data_ptr(), rows, cols, ptr(sc), _lib.stream_ptr(),
                   launches=1 if sc is None else 2)
        return sc

    # ---------------------------------------------------------------------------------------------- forward
    @on_model_device
    def forward_train(self, noisy: torch.Tensor):
        m = self.model
        self.ensure_packed()
        pk, meta, lib = self.pk, self.meta, self.lib
        if m.glu_activation != "Sigmoid":
            raise NotImplementedError("cleanumamba_b200 backward: glu_activation='Sigmoid' only")
        B, _, L = noisy.shape
        D = meta["D"]
        st = _lib.stream_ptr
        S: dict = dict(B=B, L=L)                              # saved tensors
        x = noisy
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = noisy.to(torch.float32).contiguous()
        std = None
        if m.normalize_input:
            std = self.new(B)
            self._call("wave_normalize", lib.cum_wave_normalize_fwd, x.data_ptr(), std.data_ptr(), B, L, st())
            if x is not noisy:
                noisy.copy_(x)
        S["x"], S["std"] = x, std
        Ls = [m.valid_length(L)]
        for _ in range(D):
            Ls.append((Ls[-1] - 4) // 2 + 1)
        S["Ls"] = Ls
        S["y"], S["Z"], S["skip"] = [], [], []
        prev = None
        fuse = self.fused_forward and self.math != _lib.MATH_FP32
        for i, e in enumerate(meta["enc"]):
            rows = B * Ls[i + 1]
            y = self.new(rows, e["Hc_p"])
            if i == 0:
                self._call("conv_in", lib.cum_conv_in_fwd, x.data_ptr(), L, B, L, pk["enc0.w"].data_ptr(),
                           pk["enc0.b"].data_ptr(), y.data_ptr(), Ls[1], e["Hc_p"], 4, 2, 0, 0, 0, st())
            else:
                cp = e["Cin_p"]
                self.gemm(prev, 0, Ls[i] * cp, 2 * cp, Ls[i] // 2, 2 * cp, f"enc{i}.w", pk[f"enc{i}.b"],
                          y, 0, Ls[i + 1] * e["Hc_p"], e["Hc_p"], Ls[i + 1], e["Hc_p"], B, EPI_RELU, taps=2, shifts=(0, 1))
            if fuse:        # gate in the epilogue, pre-activation saved through the second output
                Z = self.new(rows, 2 * e["Ho_p"])
                prev = self.dense(y, rows, e["Hc_p"], f"enc{i}.wg", pk[f"enc{i}.bg"], 2 * e["Ho_p"], epi=EPI_GLU["Sigmoid"], aux=Z)
            else:
                Z = self.dense(y, rows, e["Hc_p"], f"enc{i}.wg", pk[f"enc{i}.bg"], 2 * e["Ho_p"])
                prev = self.new(rows, e["Ho_p"])
                self._call("glu_fwd", lib.cum_glu_fwd, Z.data_ptr(), 0, prev.data_ptr(), rows, e["Ho_p"], st())
            S["y"].append(y); S["Z"].append(Z); S["skip"].append(prev)

        T = Ls[D]
        rows = B * T
        dm, dm_p = meta["dm"], meta["dm_p"]
        cb_p = meta["enc"][-1]["Ho_p"]
        h = self.dense(prev, rows, cb_p, "t1.w", pk["t1.b"], dm_p)
        S["mamba"] = []
        res = None
        nchunks = (T + 15) // 16
        for l, mm in enumerate(meta["mamba"]):
            di_p, N_p, R_p = mm["di_p"], mm["N_p"], mm["R_p"]
            res_out, hn = self.new(rows, dm_p), self.new(rows, dm_p)
            self.ln(h, res, res_out, hn, pk[f"m{l}.g"], pk[f"m{l}.be"], mm["eps"], rows, dm, dm_p)
            res = res_out
            xz = self.dense(hn, rows, dm_p, f"m{l}.in", None, 2 * di_p)
            xc = self.new(rows, di_p)
            self._call("dwconv_silu", lib.cum_dwconv_silu_fwd, xz.data_ptr(), T * 2 * di_p, 2 * di_p,
                       pk[f"m{l}.cw"].data_ptr(), pk[f"m{l}.cb"].data_ptr(), xc.data_ptr(), 0, 0, B, T, di_p, mm["W"], st())
            xdbl = self.dense(xc, rows, di_p, f"m{l}.xp", None, R_p + 2 * N_p)
            dt = self.dense(xdbl, rows, R_p, f"m{l}.dtw", None, di_p, a_rs=R_p + 2 * N_p)
            y = self.new(rows, di_p)
            ck = self.new(B, nchunks, N_p, di_p)          # chunk-start states, channel fastest (cum_scan_desc.h_ckpt)
            self.scan(xc, dt, xz, xdbl, y, l, mm, B, T, h_ckpt=ck)
            h = self.dense(y, rows, di_p, f"m{l}.out", None, dm_p)
            S["mamba"].append(dict(res=res, hn=hn, xz=xz, xc=xc, xdbl=xdbl, dt=dt, y=y, ck=ck))
        res_f, hn_f = self.new(rows, dm_p), self.new(rows, dm_p)
        self.ln(h, res, res_f, hn_f, pk["nf.g"], pk["nf.be"], meta["eps"], rows, dm, dm_p)
        S["res_f"], S["hn_f"] = res_f, hn_f
        xcur = self.dense(hn_f, rows, dm_p, "t2.w", pk["t2.b"], cb_p, addend=S["skip"][D - 1])

        S["xin"], S["Zd"], S["g"], S["r"] = [], [], [], []
        Tj = T
        out = None
        for j, d in enumerate(meta["dec"]):
            S["xin"].append(xcur)
            if fuse:
                Zd = self.new(B * Tj, 2 * d["Hg_p"])
                g = self.dense(xcur, B * Tj, d["Cin_p"], f"dec{j}.wg", pk[f"dec{j}.bg"], 2 * d["Hg_p"], epi=EPI_GLU["Sigmoid"], aux=Zd)
            else:
                Zd = self.dense(xcur, B * Tj, d["Cin_p"], f"dec{j}.wg", pk[f"dec{j}.bg"], 2 * d["Hg_p"])
                g = self.new(B * Tj, d["Hg_p"])
                self._call("glu_fwd", lib.cum_glu_fwd, Zd.data_ptr(), 0, g.data_ptr(), B * Tj, d["Hg_p"], st())
            S["Zd"].append(Zd); S["g"].append(g)
            if j < D - 1:
                co = d["Co_p"]
                To = 2 * Tj + 2
                r = self.new(B * To, co)
                nxt = self.new(B * To, co)
                if fuse:    # skip add in the epilogue; the pre-add ReLU output (the mask of the backward) leaves through the second output
                    skip = S["skip"][D - 2 - j]
                    self.gemm(g, 0, Tj * d["Hg_p"], d["Hg_p"], Tj, d["Hg_p"], f"dec{j}.w", pk[f"dec{j}.b"],
                              nxt, 0, To * co, 2 * co, Tj + 1, 2 * co, B, EPI_RELU, taps=2, shifts=(0, -1),
                              addend=skip, add_bs=To * co, add_rs=2 * co, aux=r, aux_bs=To * co, aux_rs=2 * co)
                else:
                    self.gemm(g, 0, Tj * d["Hg_p"], d["Hg_p"], Tj, d["Hg_p"], f"dec{j}.w", pk[f"dec{j}.b"],
                              r, 0, To * co, 2 * co, Tj + 1, 2 * co, B, EPI_RELU, taps=2, shifts=(0, -1))
                    self._call("add", lib.cum_add_fwd, r.data_ptr(), S["skip"][D - 2 - j].data_ptr(), nxt.data_ptr(), B * To * co, st())
                S["r"].append(r)
                xcur, Tj = nxt, To
            else:
                length = L if m.normalize_input else Ls[0]
                out = self.new(B, 1, length)
                self._call("convt_out", lib.cum_convt_out_fwd, g.data_ptr(), B, Tj, d["Hg_p"], pk[f"dec{j}.w"].data_ptr(),
                           meta["out_bias"], ptr(std), length, out.data_ptr(), length, 0, length, 4, 2, st())
                S["T_last"], S["length"] = Tj, length
        return out, S

    # ---------------------------------------------------------------------------------------------- backward
    @on_model_device
    def backward(self, S: dict, dout: torch.Tensor, sync=None) -> Dict[str, torch.Tensor]:
        """dout: (B, 1, length) -> flat gradient views ``self.gk`` (packed layout), also returned.  ``sync`` (a
        distributed.GradSync) starts an asynchronous all-reduce of each gradient bucket as soon as it is complete."""
        def bucket_done(name):
            if sync is not None:
                lo, hi = self.buckets[name]
                sync.reduce(self.gflat[lo:hi])
        m, pk, meta, lib = self.model, self.pk, self.meta, self.lib
        st = _lib.stream_ptr
        B, L, D, Ls = S["B"], S["L"], meta["D"], S["Ls"]
        self.gflat.zero_()
        gk = self.gk
        dout = dout.to(torch.float32).contiguous()
        dskip: List[torch.Tensor] = [None] * D

        # ---- last decoder level: transposed conv to the waveform
        j = D - 1
        d = meta["dec"][j]
        Tj = S["T_last"]
        dg = self.new(B * Tj, d["Hg_p"])
        self._call("convt_out_bwd", lib.cum_convt_out_bwd, S["g"][j].data_ptr(), B, Tj, d["Hg_p"], pk[f"dec{j}.w"].data_ptr(),
                   ptr(S["std"]), dout.data_ptr(), S["length"], S["length"], dg.data_ptr(), gk[f"dec{j}.w"].data_ptr(),
                   gk["out_bias"].data_ptr(), 4, 2, st())
        # ---- decoder levels, deepest last
        for j in range(D - 1, -1, -1):
            d = meta["dec"][j]
            rows = B * Tj
            Zd, hg, cin = S["Zd"][j], d["Hg_p"], d["Cin_p"]
            sc = self.scale_slot()
            self._call("glu_bwd", lib.cum_glu_bwd, Zd.data_ptr(), dg.data_ptr(), Zd.data_ptr(), gk[f"dec{j}.bg"].data_ptr(),
                       rows, hg, ptr(sc), st(), launches=1 if sc is None else 2)                 # dZ in place (+ its scale)
            self.wgrad(Zd, 0, 2 * hg, S["xin"][j], 0, 0, cin, rows, f"dec{j}.wg", rows, 2 * hg, cin, 1, scale=sc)
            fuse_relu = j > 0 and self.fused_relu_bwd and self.fused_forward and self.math != _lib.MATH_FP32
            if fuse_relu:
                # x_in = relu(convT(prev)) + skip: dskip = dx (unmasked, second output), d(convT output) = dx where r > 0, written over r
                r = S["r"][j - 1]
                dx = self.new(rows, cin)
                self.dense_T(Zd, rows, 2 * hg, f"dec{j}.wg", cin, scale=sc, relu_mask=r, unmasked=dx, out=r)
            else:
                dx = self.dense_T(Zd, rows, 2 * hg, f"dec{j}.wg", cin, scale=sc)           # gradient of this level's input
            if j > 0:
                lvl = D - 1 - j                                               # x_in = relu-convT(prev) + skip[lvl]
                dskip[lvl] = dx
                dp = meta["dec"][j - 1]
                co, hgp, Tp = dp["Co_p"], dp["Hg_p"], (Tj - 2) // 2
                r = S["r"][j - 1]
                if fuse_relu:
                    sc = self.colsum_scale(r, f"dec{j-1}.b", B * (Tp + 1), 2 * co)
                else:
                    sc = self.scale_slot()
                    self._call("relu_bwd", lib.cum_relu_bwd, r.data_ptr(), dx.data_ptr(), r.data_ptr(), gk[f"dec{j-1}.b"].data_ptr(),
                               B * (Tp + 1), 2 * co, ptr(sc), st(), launches=1 if sc is None else 2)     # dZ (B, Tp+1, 2co) in place of r
                self.wgrad(r, (Tp + 1) * 2 * co, 2 * co, S["g"][j - 1], 0, Tp * hgp, hgp, Tp, f"dec{j-1}.w", Tp + 1, 2 * co,
                           hgp, B, taps=2, shifts=(0, -1), scale=sc)
                dg = self.new(B * Tp, hgp)
                self.gemm(r, 0, (Tp + 1) * 2 * co, 2 * co, Tp + 1, 2 * co, f"dec{j-1}.wT", None, dg, 0, Tp * hgp, hgp, Tp, hgp, B,
                          EPI_NONE, taps=2, shifts=(0, 1), a_scale=sc)
                Tj = Tp
            else:
                dskip[D - 1] = dx                                             # x_in = tsfm_conv2(hn_f) + skip[D-1]
        bucket_done("decoder")
        # ---- tsfm_conv2
        T = Ls[D]
        rows = B * T
        dm, dm_p, cb_p = meta["dm"], meta["dm_p"], meta["enc"][-1]["Ho_p"]
        dx0 = dskip[D - 1]
        sc = self.colsum_scale(dx0, "t2.b", rows, cb_p)
        self.wgrad(dx0, 0, cb_p, S["hn_f"], 0, 0, dm_p, rows, "t2.w", rows, cb_p, dm_p, 1, scale=sc)
        dhn = self.dense_T(dx0, rows, cb_p, "t2.w", dm_p, scale=sc)
        # ---- final norm, Mamba layers in reverse
        dres = self.new(rows, dm_p)
        self._call("ln_bwd", lib.cum_ln_residual_bwd, S["res_f"].data_ptr(), dhn.data_ptr(), 0, pk["nf.g"].data_ptr(),
                   dres.data_ptr(), gk["nf.g"].data_ptr(), gk["nf.be"].data_ptr(), meta["eps"], rows, dm, dm_p, st())
        for l in range(len(meta["mamba"]) - 1, -1, -1):
            mm, sv = meta["mamba"][l], S["mamba"][l]
            di_p, N_p, R_p = mm["di_p"], mm["N_p"], mm["R_p"]
            ld = R_p + 2 * N_p
            dh = dres                                                          # grad of the mixer output == grad of res_{l+1}
            sc = self.grad_scale(dh, rows, dm_p)
            self.wgrad(dh, 0, dm_p, sv["y"], 0, 0, di_p, rows, f"m{l}.out", rows, dm_p, di_p, 1, scale=sc)
            dy = self.dense_T(dh, rows, dm_p, f"m{l}.out", di_p, scale=sc)
            dxz = self.new(rows, 2 * di_p)
            dxdbl = self.zeros(rows, ld)
            du, ddt = self.new(rows, di_p), self.new(rows, di_p)
            sb = ScanBwdDesc()
            self.fill_scan(sb.fwd, sv["xc"], sv["dt"], sv["xz"], sv["xdbl"], None, l, mm, B, T)
            sb.h_ckpt = sv["ck"].data_ptr()
            sb.dout, sb.dout_bs, sb.dout_rs = dy.data_ptr(), T * di_p, di_p
            sb.du, sb.du_bs, sb.du_rs = du.data_ptr(), T * di_p, di_p
            sb.ddelta, sb.ddl_bs, sb.ddl_rs = ddt.data_ptr(), T * di_p, di_p
            sb.dz, sb.dz_bs, sb.dz_rs = dxz.data_ptr() + 4 * di_p, T * 2 * di_p, 2 * di_p
            sb.dB, sb.dB_bs, sb.dB_rs = dxdbl.data_ptr() + 4 * R_p, T * ld, ld
            sb.dC, sb.dC_bs, sb.dC_rs = dxdbl.data_ptr() + 4 * (R_p + N_p), T * ld, ld
            sb.dA_log, sb.dD, sb.ddelta_bias = gk[f"m{l}.a2"].data_ptr(), gk[f"m{l}.D"].data_ptr(), gk[f"m{l}.dtb"].data_ptr()
            self._call("selective_scan_bwd", lib.cum_selective_scan_bwd, C.byref(sb), st())
            # dt_proj: delta = x_dbl[:, :R] @ W_dt^T
            sc = self.grad_scale(ddt, rows, di_p)
            self.wgrad(ddt, 0, di_p, sv["xdbl"], 0, 0, ld, rows, f"m{l}.dtw", rows, di_p, R_p, 1, scale=sc)
            self.dense_T(ddt, rows, di_p, f"m{l}.dtw", R_p, out=dxdbl, c_rs=ld, scale=sc)
            # x_proj
            sc = self.grad_scale(dxdbl, rows, ld)
            self.wgrad(dxdbl, 0, ld, sv["xc"], 0, 0, di_p, rows, f"m{l}.xp", rows, ld, di_p, 1, scale=sc)
            dxc = self.dense_T(dxdbl, rows, ld, f"m{l}.xp", di_p, addend=du, scale=sc)
            # depthwise conv + SiLU (x half of xz)
            self._call("dwconv_silu_bwd", lib.cum_dwconv_silu_bwd, sv["xz"].data_ptr(), T * 2 * di_p, 2 * di_p,
                       pk[f"m{l}.cw"].data_ptr(), pk[f"m{l}.cb"].data_ptr(), dxc.data_ptr(), dxz.data_ptr(), T * 2 * di_p,
                       2 * di_p, gk[f"m{l}.cw"].data_ptr(), gk[f"m{l}.cb"].data_ptr(), B, T, di_p, mm["W"], st())
            # in_proj
            sc = self.grad_scale(dxz, rows, 2 * di_p)
            self.wgrad(dxz, 0, 2 * di_p, sv["hn"], 0, 0, dm_p, rows, f"m{l}.in", rows, 2 * di_p, dm_p, 1, scale=sc)
            dhn = self.dense_T(dxz, rows, 2 * di_p, f"m{l}.in", dm_p, scale=sc)
            # pre-norm + residual stream
            dres_new = self.new(rows, dm_p)
            self._call("ln_bwd", lib.cum_ln_residual_bwd, sv["res"].data_ptr(), dhn.data_ptr(), dres.data_ptr(),
                       pk[f"m{l}.g"].data_ptr(), dres_new.data_ptr(), gk[f"m{l}.g"].data_ptr(), gk[f"m{l}.be"].data_ptr(),
                       mm["eps"], rows, dm, dm_p, st())
            dres = dres_new
        # ---- tsfm_conv1 (its output is res_0)
        sc = self.colsum_scale(dres, "t1.b", rows, dm_p)
        self.wgrad(dres, 0, dm_p, S["skip"][D - 1], 0, 0, cb_p, rows, "t1.w", rows, dm_p, cb_p, 1, scale=sc)
        dskip[D - 1] = self.dense_T(dres, rows, dm_p, "t1.w", cb_p, addend=dskip[D - 1], scale=sc)
        bucket_done("bottleneck")
        # ---- encoder levels in reverse
        for i in range(D - 1, -1, -1):
            e = meta["enc"][i]
            rows = B * Ls[i + 1]
            Z, ho, hc = S["Z"][i], e["Ho_p"], e["Hc_p"]
            sc = self.scale_slot()
            self._call("glu_bwd", lib.cum_glu_bwd, Z.data_ptr(), dskip[i].data_ptr(), Z.data_ptr(), gk[f"enc{i}.bg"].data_ptr(),
                       rows, ho, ptr(sc), st(), launches=1 if sc is None else 2)
            self.wgrad(Z, 0, 2 * ho, S["y"][i], 0, 0, hc, rows, f"enc{i}.wg", rows, 2 * ho, hc, 1, scale=sc)
            y = S["y"][i]
            fuse_relu = i > 0 and self.fused_relu_bwd and self.fused_forward and self.math != _lib.MATH_FP32
            if fuse_relu:       # ReLU backward in the epilogue of the data-gradient GEMM: dZ = dy where y > 0, written over y
                self.dense_T(Z, rows, 2 * ho, f"enc{i}.wg", hc, scale=sc, relu_mask=y, out=y)
                dy = None
            else:
                dy = self.dense_T(Z, rows, 2 * ho, f"enc{i}.wg", hc, scale=sc)
            if i > 0:
                cp = e["Cin_p"]
                if fuse_relu:
                    sc = self.colsum_scale(y, f"enc{i}.b", rows, hc)
                else:
                    sc = self.scale_slot()
                    self._call("relu_bwd", lib.cum_relu_bwd, y.data_ptr(), dy.data_ptr(), y.data_ptr(), gk[f"enc{i}.b"].data_ptr(),
                               rows, hc, ptr(sc), st(), launches=1 if sc is None else 2)             # dZ in place of y
                src = S["skip"][i - 1]
                self.wgrad(y, Ls[i + 1] * hc, hc, src, 0, Ls[i] * cp, 2 * cp, Ls[i] // 2, f"enc{i}.w", Ls[i + 1], hc, 2 * cp, B,
                           taps=2, shifts=(0, 1), scale=sc)
                dprev = self.new(B * Ls[i], cp)
                self.gemm(y, 0, Ls[i + 1] * hc, hc, Ls[i + 1], hc, f"enc{i}.wT", None, dprev, 0, Ls[i] * cp, 2 * cp, Ls[i] // 2,
                          2 * cp, B, EPI_NONE, taps=2, shifts=(0, -1), addend=dskip[i - 1], add_bs=Ls[i] * cp, add_rs=2 * cp, a_scale=sc)
                dskip[i - 1] = dprev
                if i == self.enc_split_level:
                    bucket_done("encoder_deep")
            else:
                self._call("conv_in_bwd", lib.cum_conv_in_bwd, S["x"].data_ptr(), L, B, L, y.data_ptr(), dy.data_ptr(),
                           gk["enc0.w"].data_ptr(), gk["enc0.b"].data_ptr(), Ls[1], hc, 4, 2, st())
        bucket_done("encoder_outer")
        return gk

    # ---------------------------------------------------------------------------------------------- unpack
    @torch.no_grad()
    def unpack_grads(self) -> Dict[str, torch.Tensor]:
        """Packed-layout gradients -> {state_dict key: gradient with the parameter's shape}."""
        m, meta, gk = self.model, self.meta, self.gk
        D = meta["D"]
        out: Dict[str, torch.Tensor] = {}

        def deinterleave(gw, gb, H, K):
            w = torch.cat([gw[0:2 * H:2, :K], gw[1:2 * H:2, :K]], 0)
            b = torch.cat([gb[0:2 * H:2], gb[1:2 * H:2]], 0)
            return w, b
        for i, e in enumerate(meta["enc"]):
            Hc = e["Hc"]
            if i == 0:
                out["encoder.0.0.weight"] = gk["enc0.w"][:, :Hc].t()[:, None, :].clone()
            else:
                cp = e["Cin_p"]
                Cin = m.encoder[i][0].weight.shape[1]
                g = gk[f"enc{i}.w"]
                w = g.new_zeros(Hc, Cin, 4)
                for s in range(2):
                    for jj in range(2):
                        w[:, :, 2 * s + jj] = g[s, :Hc, jj * cp: jj * cp + Cin]
                out[f"encoder.{i}.0.weight"] = w
            out[f"encoder.{i}.0.bias"] = gk[f"enc{i}.b"][:Hc].clone()
            w, b = deinterleave(gk[f"enc{i}.wg"], gk[f"enc{i}.bg"], e["Ho"], Hc)
            out[f"encoder.{i}.2.weight"], out[f"encoder.{i}.2.bias"] = w[:, :, None], b          # torch.cat results: fresh storage
        dm = meta["dm"]
        C_b = meta["enc"][-1]["Ho"]
        # .clone(), never .contiguous(): when no padding is cut away (dm % 8 == 0, every full-size model) .contiguous() is a VIEW of
        # the persistent gradient buffer, autograd would adopt it as param.grad and the next backward would overwrite / double it
        out["tsfm_conv1.weight"] = gk["t1.w"][:dm, :C_b, None].clone()
        out["tsfm_conv1.bias"] = gk["t1.b"][:dm].clone()
        for l, mm in enumerate(meta["mamba"]):
            di, di_p, N, N_p, R, R_p = mm["di"], mm["di_p"], mm["N"], mm["N_p"], mm["R"], mm["R_p"]
            p = f"tsfm_Mamba_layers.{l}."
            gin = gk[f"m{l}.in"]
            out[p + "mixer.in_proj.weight"] = torch.cat([gin[:di, :dm], gin[di_p: di_p + di, :dm]], 0)
            out[p + "mixer.conv1d.weight"] = gk[f"m{l}.cw"][:, :di].t()[:, None, :].clone()
            out[p + "mixer.conv1d.bias"] = gk[f"m{l}.cb"][:di].clone()
            gx = gk[f"m{l}.xp"]
            out[p + "mixer.x_proj.weight"] = torch.cat([gx[:R, :di], gx[R_p: R_p + N, :di], gx[R_p + N_p: R_p + N_p + N, :di]], 0)
            out[p + "mixer.dt_proj.weight"] = gk[f"m{l}.dtw"][:di, :R].clone()
            out[p + "mixer.dt_proj.bias"] = gk[f"m{l}.dtb"][:di].clone()
            out[p + "mixer.A_log"] = gk[f"m{l}.a2"][:di, :N].clone()          # the scan accumulates dA_log here
            out[p + "mixer.D"] = gk[f"m{l}.D"][:di].clone()
            out[p + "mixer.out_proj.weight"] = gk[f"m{l}.out"][:dm, :di].clone()
            out[p + "norm.weight"], out[p + "norm.bias"] = gk[f"m{l}.g"][:dm].clone(), gk[f"m{l}.be"][:dm].clone()
        out["norm_f.weight"], out["norm_f.bias"] = gk["nf.g"][:dm].clone(), gk["nf.be"][:dm].clone()
        out["tsfm_conv2.weight"] = gk["t2.w"][:C_b, :dm, None].clone()
        out["tsfm_conv2.bias"] = gk["t2.b"][:C_b].clone()
        for j, d in enumerate(meta["dec"]):
            Cin = m.decoder[j][0].weight.shape[1]
            w, b = deinterleave(gk[f"dec{j}.wg"], gk[f"dec{j}.bg"], d["Hg"], Cin)
            out[f"decoder.{j}.0.weight"], out[f"decoder.{j}.0.bias"] = w[:, :, None], b
            Hg, Co, Co_p = d["Hg"], d["Co"], d["Co_p"]
            if j == D - 1:
                out[f"decoder.{j}.2.weight"] = gk[f"dec{j}.w"][:, :Hg].t()[:, None, :].clone()
                out[f"decoder.{j}.2.bias"] = gk["out_bias"][:1].clone()
            else:
                g = gk[f"dec{j}.w"]
                wt = g.new_zeros(Hg, Co, 4)
                for s in range(2):
                    for par in range(2):
                        wt[:, :, 2 * s + par] = g[s, par * Co_p: par * Co_p + Co, :Hg].t()
                out[f"decoder.{j}.2.weight"] = wt
                gb = gk[f"dec{j}.b"]
                out[f"decoder.{j}.2.bias"] = gb[:Co] + gb[Co_p: Co_p + Co]
        return out
