"""CPU suite, round-2 additions: the new product entry points fail loudly without a GPU (no fallback), the reference staging used
by the CPU arms of bench.py, and the DFT basis of the fused STFT loss against torch.stft."""
import math
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_new_entry_points_have_no_cpu_fallback():
    from cleanumamba_b200.fused_loss import FusedMultiResolutionSTFTLoss
    from cleanumamba_b200.importance import channel_importances
    from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG
    x = torch.randn(2, 4000)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedMultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        channel_importances(torch.randn(8, 4), torch.randn(8, 4))
    with pytest.raises(NotImplementedError):
        FusedMultiResolutionSTFTLoss(band="high")


def test_reference_staging_is_byte_exact_and_git_ignored():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stage_reference
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not mounted")
    assert stage_reference.stage() is True
    for rel in stage_reference.FILES:
        assert open(os.path.join(stage_reference.DST, rel), "rb").read() == open(os.path.join("/root/reference", rel), "rb").read(), rel
    assert "baseline/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()
    gi = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gi) or "baseline" not in open(gi).read()      # the staged copy must travel to the GPU box


def test_windowed_dft_basis_reproduces_torch_stft():
    """fused_loss folds the hann window into a (2F, win_length) cos / -sin basis and contracts only the win_length non-zero taps of
    the centred n_fft window; frames are cut with torch.stft's reflect padding.  Check that algebra on the CPU (fp64) against
    torch.stft for the three shipped resolutions (the GEMM itself is the GPU test's business)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6000, generator=g, dtype=torch.float64)
    for n_fft, hop, win in ((512, 50, 240), (1024, 120, 600), (2048, 240, 1200)):
        bins = n_fft // 2 + 1
        w = torch.hann_window(win, dtype=torch.float64)
        n = torch.arange(win, dtype=torch.float64) + (n_fft - win) // 2
        ang = 2.0 * math.pi * torch.arange(bins, dtype=torch.float64)[:, None] * n[None, :] / n_fft
        basis = torch.zeros(2 * bins, win, dtype=torch.float64)
        basis[0::2], basis[1::2] = torch.cos(ang) * w, -torch.sin(ang) * w
        nf = 1 + x.shape[1] // hop
        first = (n_fft - win) // 2 - n_fft // 2
        idx = torch.arange(nf)[:, None] * hop + first + torch.arange(win)[None, :]
        idx = idx.abs()
        idx = torch.where(idx >= x.shape[1], 2 * (x.shape[1] - 1) - idx, idx)       # reflect padding by index arithmetic
        frames = x[:, idx]                                                          # (B, nf, win)
        S = frames @ basis.t()                                                      # (B, nf, 2 bins)
        ref = torch.stft(x, n_fft, hop, win, w, return_complex=True).transpose(1, 2)    # (B, nf, bins)
        assert (S[..., 0::2] - ref.real).abs().max() < 1e-9 and (S[..., 1::2] - ref.imag).abs().max() < 1e-9
