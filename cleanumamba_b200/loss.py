"""Training loss of the reference, kept in PyTorch as BASELINE.json's north_star specifies (cuFFT via torch.stft).

Restates /root/reference/src/util/util.py:215-327 (``loss_fn``: L1 / L2 reconstruction + multi-resolution STFT term;
the knowledge-distillation and cross-entropy branches are outside the hot path) and
/root/reference/src/util/stft_loss.py:16-184 (spectral convergence + log-magnitude L1 over several STFT resolutions).
Default resolutions / weights are those of configs/config.json:23-36."""
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

DEFAULT_STFT_CONFIG = dict(sc_lambda=0.5, mag_lambda=0.5, band="full", hop_sizes=[50, 120, 240],
                           win_lengths=[240, 600, 1200], fft_sizes=[512, 1024, 2048])


def stft_magnitude(x: torch.Tensor, fft_size: int, hop: int, win_length: int, window: torch.Tensor) -> torch.Tensor:
    """(B, T) -> (B, frames, fft_size//2+1) magnitudes, floored at sqrt(1e-7) like the reference (stft_loss.py:16-35)."""
    spec = torch.stft(x, fft_size, hop, win_length, window, return_complex=True)
    power = spec.real.square() + spec.imag.square()
    return torch.sqrt(torch.clamp(power, min=1e-7)).transpose(2, 1)


class MultiResolutionSTFTLoss(torch.nn.Module):
    def __init__(self, fft_sizes: Sequence[int] = (1024, 2048, 512), hop_sizes: Sequence[int] = (120, 240, 50),
                 win_lengths: Sequence[int] = (600, 1200, 240), window: str = "hann_window", sc_lambda: float = 0.1,
                 mag_lambda: float = 0.1, band: str = "full"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        if band not in ("full", "high"):
            raise NotImplementedError(band)
        self.resolutions = list(zip(fft_sizes, hop_sizes, win_lengths))
        self.sc_lambda, self.mag_lambda, self.band = sc_lambda, mag_lambda, band
        for i, (_, _, wl) in enumerate(self.resolutions):
            self.register_buffer(f"window_{i}", getattr(torch, window)(wl))

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if x.dim() == 3:
            x, y = x.reshape(-1, x.size(2)), y.reshape(-1, y.size(2))
        sc_total, mag_total = 0.0, 0.0
        for i, (fs, hop, wl) in enumerate(self.resolutions):
            win = getattr(self, f"window_{i}")
            xm, ym = stft_magnitude(x, fs, hop, wl, win), stft_magnitude(y, fs, hop, wl, win)
            if self.band == "high":          # the reference slices dim 1 (frames) here (stft_loss.py:117-119); kept as is
                half = xm.shape[1] // 2
                xm, ym = xm[:, half:, :], ym[:, half:, :]
            sc_total = sc_total + torch.linalg.norm(ym - xm) / torch.linalg.norm(ym)
            mag_total = mag_total + F.l1_loss(torch.log(ym), torch.log(xm))
        n = len(self.resolutions)
        return sc_total * self.sc_lambda / n, mag_total * self.mag_lambda / n


def loss_fn(net, X, ell_p: int = 1, ell_p_lambda: float = 1, stft_lambda: float = 1, mrstftloss=None,
            **unused) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """X = (clean, noisy), both (B, 1, L).  Returns (loss, parts) like the reference's ``loss_fn``."""
    clean, noisy = X
    denoised = net(noisy)
    if ell_p == 1:
        ae = F.l1_loss(denoised, clean)
    elif ell_p == 2:
        ae = F.mse_loss(denoised, clean)
    else:
        raise NotImplementedError(ell_p)
    loss = ae * ell_p_lambda
    parts = {"reconstruct": ae.detach() * ell_p_lambda}
    if stft_lambda > 0:
        if mrstftloss is None:
            mrstftloss = MultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG).to(denoised.device)
        sc, mag = mrstftloss(denoised.squeeze(1), clean.squeeze(1))
        loss = loss + (sc + mag) * stft_lambda
        parts["stft_sc"], parts["stft_mag"] = sc.detach() * stft_lambda, mag.detach() * stft_lambda
    return loss, parts
