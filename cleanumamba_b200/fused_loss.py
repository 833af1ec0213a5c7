"""Fused multi-resolution STFT loss on B200 -- drop-in for ``MultiResolutionSTFTLoss``
(/root/reference/src/util/stft_loss.py:130-184; called from ``loss_fn``, src/util/util.py:322) including its backward.

Per resolution the reference runs torch.stft (cuFFT) on both signals and ~10 elementwise / reduction kernels, and autograd
replays them backwards.  Here one resolution is

    frames  (reflect padding by index arithmetic)                      cum_stft_frames_fwd
    S = frames . basis^T   windowed DFT as a tcgen05 GEMM, f16x3        cum_gemm_bias_act_fwd
        (hann window folded into the (2F, win_length) cos / -sin basis: only the win_length non-zero taps are contracted)
    three sums over both spectrograms                                   cum_stft_loss_reduce_fwd
and the backward
    dL/dS of the predicted signal                                       cum_stft_loss_bwd
    dframes = dS . basis   (tf32x3: gradients need fp32 range)          cum_gemm_bias_act_fwd
    overlap-add onto the waveform                                       cum_stft_overlap_add

``loss.MultiResolutionSTFTLoss`` (PyTorch, pinned to the reference run live) stays as the checker.  CUDA only, no fallback."""
import ctypes as C
import math
from typing import Sequence, Tuple

import torch

from . import _lib
from ._lib import GemmDesc, check


def _p8(n):
    return (n + 7) // 8 * 8


class _Resolution:
    """Device-resident DFT basis of one (n_fft, hop, win_length) resolution, split for the two GEMM modes."""

    def __init__(self, n_fft, hop, win, window, device, lib):
        self.n_fft, self.hop, self.win, self.bins = n_fft, hop, win, n_fft // 2 + 1
        self.n_pad = _p8(2 * self.bins)
        if win % 8:
            raise NotImplementedError(f"win_length={win}: the DFT GEMM needs a multiple of 8")
        w = getattr(torch, window)(win, dtype=torch.float64)
        n = torch.arange(win, dtype=torch.float64) + (n_fft - win) // 2          # the window sits centred inside the n_fft frame
        ang = 2.0 * math.pi * torch.arange(self.bins, dtype=torch.float64)[:, None] * n[None, :] / n_fft
        basis = torch.zeros(self.n_pad, win, dtype=torch.float64)
        basis[0:2 * self.bins:2] = torch.cos(ang) * w
        basis[1:2 * self.bins:2] = -torch.sin(ang) * w
        self.basis = basis.to(torch.float32).to(device).contiguous()
        e = int(math.floor(math.log2(8.0 / float(self.basis.abs().max()))))
        self.scale_inv = float(2.0 ** -e)
        self.b_hi = torch.empty(self.basis.shape, dtype=torch.float16, device=device)
        self.b_lo = torch.empty_like(self.b_hi)
        check(lib.cum_split_f16(self.basis.data_ptr(), self.b_hi.data_ptr(), self.b_lo.data_ptr(), self.basis.numel(), float(2.0 ** e),
                                _lib.stream_ptr()), "cum_split_f16")
        self.basis_t = self.basis.t().contiguous()                                # (win, n_pad): weights of the backward GEMM
        self.bt_hi, self.bt_lo = torch.empty_like(self.basis_t), torch.empty_like(self.basis_t)
        check(lib.cum_split_tf32(self.basis_t.data_ptr(), self.bt_hi.data_ptr(), self.bt_lo.data_ptr(), self.basis_t.numel(),
                                 _lib.stream_ptr()), "cum_split_tf32")


def _gemm(lib, a, rows, k, w_hi, w_lo, ldw, c, n, math_mode, acc_scale=1.0):
    d = GemmDesc()
    d.a, d.a_batch_stride, d.a_row_stride, d.a_rows, d.k, d.taps = a.data_ptr(), 0, k, rows, k, 1
    d.w, d.w_lo, d.ldw, d.bias = w_hi.data_ptr(), w_lo.data_ptr(), ldw, 0
    d.c, d.c_batch_stride, d.c_row_stride, d.m, d.n, d.batch = c.data_ptr(), 0, n, rows, n, 1
    d.epilogue, d.math, d.acc_scale = _lib.EPI_NONE, math_mode, acc_scale
    check(lib.cum_gemm_bias_act_fwd(C.byref(d), _lib.stream_ptr()), "cum_gemm_bias_act_fwd")


class _MRSTFTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, mod):
        lib, dev = mod._lib(x.device), x.device
        B, T = x.shape
        xc, yc = x.detach().float().contiguous(), y.detach().float().contiguous()
        sums = torch.zeros(len(mod.res), 3, dtype=torch.float64, device=dev)
        saved = []
        st = _lib.stream_ptr
        for i, r in enumerate(mod.res):
            nf = 1 + T // r.hop
            rows = 2 * B * nf
            frames = torch.empty(rows, r.win, dtype=torch.float32, device=dev)
            check(lib.cum_stft_frames_fwd(xc.data_ptr(), yc.data_ptr(), T, T, B, nf, r.hop, r.win, r.n_fft, frames.data_ptr(), st()),
                  "cum_stft_frames_fwd")
            S = torch.empty(rows, r.n_pad, dtype=torch.float32, device=dev)
            _gemm(lib, frames, rows, r.win, r.b_hi, r.b_lo, r.win, S, r.n_pad, _lib.MATH_F16X3, r.scale_inv)
            del frames
            check(lib.cum_stft_loss_reduce_fwd(S.data_ptr(), S[B * nf:].data_ptr(), B * nf, r.bins, r.n_pad, sums[i].data_ptr(), st()),
                  "cum_stft_loss_reduce_fwd")
            saved.append((S, nf))
        a, b2, c = sums[:, 0], sums[:, 1], sums[:, 2]
        counts = torch.tensor([B * nf * r.bins for (_, nf), r in zip(saved, mod.res)], dtype=torch.float64, device=dev)
        n = len(mod.res)
        sc = (torch.sqrt(a) / torch.sqrt(b2)).sum() * (mod.sc_lambda / n)
        mag = (c / counts).sum() * (mod.mag_lambda / n)
        ctx.mod, ctx.saved, ctx.shape, ctx.counts = mod, saved, (B, T), counts
        ctx.save_for_backward(a, b2)
        ctx.in_dtype = x.dtype
        return sc.float(), mag.float()

    @staticmethod
    def backward(ctx, g_sc, g_mag):
        mod, (B, T) = ctx.mod, ctx.shape
        a, b2 = ctx.saved_tensors
        dev = a.device
        lib = mod._lib(dev)
        n = len(mod.res)
        k0 = (g_sc.double() * (mod.sc_lambda / n) / (torch.sqrt(a).clamp_min(1e-30) * torch.sqrt(b2))).float()
        k1 = (g_mag.double() * (mod.mag_lambda / n) / ctx.counts).float()
        coef = torch.stack([k0, k1], 1).contiguous()             # (resolutions, 2)
        dx = torch.zeros(B, T, dtype=torch.float32, device=dev)
        st = _lib.stream_ptr
        for i, (r, (S, nf)) in enumerate(zip(mod.res, ctx.saved)):
            rows = B * nf
            check(lib.cum_stft_loss_bwd(S.data_ptr(), S[rows:].data_ptr(), rows, r.bins, r.n_pad, coef[i].data_ptr(), S.data_ptr(), st()),
                  "cum_stft_loss_bwd")                           # dS of the predicted half, in place
            dframes = torch.empty(rows, r.win, dtype=torch.float32, device=dev)
            _gemm(lib, S, rows, r.n_pad, r.bt_hi, r.bt_lo, r.n_pad, dframes, r.win, _lib.MATH_TF32X3)
            check(lib.cum_stft_overlap_add(dframes.data_ptr(), T, B, nf, r.hop, r.win, r.n_fft, dx.data_ptr(), T, st()),
                  "cum_stft_overlap_add")
        ctx.saved = None
        return dx.to(ctx.in_dtype), None, None


class FusedMultiResolutionSTFTLoss(torch.nn.Module):
    """Same constructor and ``forward(x, y) -> (sc_loss, mag_loss)`` as the reference's MultiResolutionSTFTLoss (x = predicted,
    y = ground truth, (B, T) or (B, 1, T)); gradients flow to ``x`` only, like the training loop uses it (util.py:322)."""

    def __init__(self, fft_sizes: Sequence[int] = (1024, 2048, 512), hop_sizes: Sequence[int] = (120, 240, 50),
                 win_lengths: Sequence[int] = (600, 1200, 240), window: str = "hann_window", sc_lambda: float = 0.1,
                 mag_lambda: float = 0.1, band: str = "full"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        if band != "full":
            raise NotImplementedError("FusedMultiResolutionSTFTLoss: band='full' only (all shipped configs)")
        self.cfg = list(zip(fft_sizes, hop_sizes, win_lengths))
        self.window, self.sc_lambda, self.mag_lambda = window, sc_lambda, mag_lambda
        self.res, self._dev = None, None

    def _lib(self, device):
        lib = _lib.init(device)
        if self.res is None or self._dev != device:
            with torch.cuda.device(device):
                self.res = [_Resolution(fs, hop, wl, self.window, device, lib) for fs, hop, wl in self.cfg]
            self._dev = device
        return lib

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if x.device.type != "cuda" or y.device != x.device:
            raise RuntimeError("FusedMultiResolutionSTFTLoss: CUDA tensors required (no CPU fallback)")
        if x.dim() == 3:
            x, y = x.reshape(-1, x.size(2)), y.reshape(-1, y.size(2))
        if x.shape != y.shape:
            raise ValueError(f"shapes differ: {tuple(x.shape)} vs {tuple(y.shape)}")
        if x.shape[1] <= max(fs for fs, _, _ in self.cfg) // 2:
            raise ValueError("signal shorter than n_fft / 2: reflect padding undefined (torch.stft raises as well)")
        with torch.cuda.device(x.device):
            return _MRSTFTFn.apply(x, y, self)
