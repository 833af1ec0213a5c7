"""Channel importances for pruning -- drop-in for the arithmetic of ``PruningModule.channel_importances``
(/root/reference/src/pruning/pruninggroup.py:160-226; consumed by ``calc_importance`` / ``get_prune_channels``,
importance.py:39-135): per channel of a prunable weight slice

    weight = sum w^2, grad = sum g^2, taylor_individual = sum |w g|, taylor_squared_individual = sum (w g)^2, taylor_group = |sum w g|

One kernel pass over the weight and its gradient (``cum_channel_importance_fwd``: 8 bytes per element, the five statistics for
every row and every column at once) replaces the reference's ~10 elementwise / reduction passes per (module, dim); the regrouping
of rows / columns into channels (heads, conv taps, the channel offset of modules that share a matrix) is a few tiny host-side
reshapes of the (5, rows) / (5, cols) results.  CUDA tensors only -- no fallback."""
from typing import Dict, Optional

import torch

from . import _lib

KEYS = ("weight", "grad", "taylor_individual", "taylor_squared_individual", "taylor_group")


@torch.no_grad()
def matrix_statistics(weight: torch.Tensor, grad: torch.Tensor):
    """(rows, ...) weight and gradient -> (stats_rows (5, rows), stats_cols (5, prod(rest))); raw sums (row 4 = signed sum w g)."""
    if weight.device.type != "cuda" or grad.device != weight.device:
        raise RuntimeError("cleanumamba_b200.importance: CUDA tensors required (no CPU fallback)")
    if weight.shape != grad.shape:
        raise ValueError(f"weight {tuple(weight.shape)} and grad {tuple(grad.shape)} differ in shape")
    w = weight.detach().reshape(weight.shape[0], -1).to(torch.float32).contiguous()
    g = grad.detach().reshape(grad.shape[0], -1).to(torch.float32).contiguous()
    rows, cols = w.shape
    out_r = torch.empty(5, rows, dtype=torch.float32, device=w.device)
    out_c = torch.empty(5, cols, dtype=torch.float32, device=w.device)
    lib = _lib.init(w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.cum_channel_importance_fwd(w.data_ptr(), g.data_ptr(), rows, cols, cols, cols, out_r.data_ptr(), out_c.data_ptr(),
                                                  _lib.stream_ptr()), "cum_channel_importance_fwd")
    return out_r, out_c


@torch.no_grad()
def channel_importances(weight: torch.Tensor, grad: Optional[torch.Tensor], dim: int = 0, channel_offset: int = 0,
                        n_channels: Optional[int] = None, n_heads: int = 1) -> Dict[str, Optional[torch.Tensor]]:
    """Same quantities, keys and grouping as the reference's ``PruningModule.channel_importances``: ``dim`` selects output (0) or
    input (1) channels of ``weight`` (1-D: a per-channel vector), ``channel_offset`` / ``n_channels`` the slice owned by the pruning
    group, ``n_heads`` consecutive rows per channel.  ``grad=None`` returns only "weight" (as the reference does without .grad)."""
    g = grad if grad is not None else torch.zeros_like(weight)
    w2 = weight if weight.dim() > 1 else weight[:, None]
    g2 = g if g.dim() > 1 else g[:, None]
    rows_stat, cols_stat = matrix_statistics(w2, g2)
    if dim == 0:
        per, n_param = rows_stat, w2[0].numel()
    elif dim == 1:
        inner = w2.shape[2:].numel() if w2.dim() > 2 else 1          # conv taps of one input channel are adjacent columns
        per = cols_stat.view(5, w2.shape[1], inner).sum(2)
        n_param = w2.shape[0] * inner
    else:
        raise ValueError("dim must be 0 or 1")
    total = per.shape[1]
    if n_channels is None:
        n_channels = (total - channel_offset) // n_heads
    sel = per[:, channel_offset: channel_offset + n_channels * n_heads].reshape(5, n_channels, n_heads).sum(2)
    out: Dict[str, Optional[torch.Tensor]] = {k: None for k in KEYS}
    out["weight"] = sel[0]
    if grad is not None:
        out["grad"], out["taylor_individual"], out["taylor_squared_individual"] = sel[1], sel[2], sel[3]
        out["taylor_group"] = sel[4].abs()
    out["n_parameters"] = n_param * n_heads
    out["act_var"] = None          # activation telemetry is gathered by the reference's forward hooks (outside the hot path)
    return out
