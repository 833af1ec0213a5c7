"""Parameter holders for the Mamba bottleneck, mirroring mamba-ssm==1.2.2 (``Mamba``, ``Block``, ``create_block``,
``_init_weights``, ``InferenceParams``; call sites /root/reference/src/network/CleanUMamba.py:12-14,172-206).

Same attribute names, parameter registration order (A_log, D, in_proj, conv1d, x_proj, dt_proj, out_proj; Block:
mixer, norm) and initialisation RNG order as the dependency, so ``state_dict`` keys/shapes match the shipped
checkpoints and a seeded constructor reproduces the reference's weights.  The arithmetic itself lives in the CUDA
library (cleanumamba_b200/csrc); these classes hold no forward of their own.
"""
import math
from dataclasses import dataclass, field
from functools import partial
from typing import Optional

import torch
import torch.nn as nn


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None):
        super().__init__()
        fk = dict(device=device, dtype=dtype)
        if bias or not conv_bias:
            raise NotImplementedError("cleanumamba_b200: Mamba(bias=False, conv_bias=True) only (reference defaults)")
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx = use_fast_path, layer_idx

        A_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(self.d_inner, 1)
        # registration order of the dependency: in_proj, conv1d, x_proj, dt_proj created first, A_log/D/out_proj after;
        # nn.Module.state_dict() lists parameters of a module before those of its children, hence A_log, D lead.
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=False, **fk)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, d_conv, groups=self.d_inner, padding=d_conv - 1, **fk)
        self.activation = "silu"
        self.act = nn.SiLU()
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False, **fk)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **fk)
        bound = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -bound, bound)
        elif dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, bound)
        else:
            raise NotImplementedError(dt_init)
        log_lo, log_hi = math.log(dt_min), math.log(dt_max)
        dt = torch.exp(torch.rand(self.d_inner, **fk) * (log_hi - log_lo) + log_lo).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))   # softplus^-1(dt)
        self.dt_proj.bias._no_reinit = True
        self.A_log = nn.Parameter(A_log)
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.d_inner, device=device))
        self.D._no_weight_decay = True
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False, **fk)

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        dev = self.out_proj.weight.device
        d = int(self.d_model * self.expand)
        return (torch.zeros(batch_size, d, self.d_conv, device=dev, dtype=dtype or self.conv1d.weight.dtype),
                torch.zeros(batch_size, d, self.d_state, device=dev, dtype=dtype or self.dt_proj.weight.dtype))

    def forward(self, hidden_states, inference_params=None):
        from . import ops
        return ops.mamba_mixer_forward(self, hidden_states, inference_params)


class Block(nn.Module):
    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False):
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.mixer = mixer_cls(dim)
        self.norm = norm_cls(dim)

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)

    def forward(self, hidden_states, residual=None, inference_params=None):
        from . import ops
        return ops.block_forward(self, hidden_states, residual, inference_params)


def create_block(d_model, ssm_cfg=None, norm_epsilon=1e-5, rms_norm=False, residual_in_fp32=False,
                 fused_add_norm=False, layer_idx=None, device=None, dtype=None):
    if rms_norm:
        raise NotImplementedError("cleanumamba_b200: rms_norm=True is not on the shipped path (LayerNorm only)")
    fk = dict(device=device, dtype=dtype)
    blk = Block(d_model, partial(Mamba, layer_idx=layer_idx, **(ssm_cfg or {}), **fk),
                norm_cls=partial(nn.LayerNorm, eps=norm_epsilon, **fk), fused_add_norm=fused_add_norm,
                residual_in_fp32=residual_in_fp32)
    blk.layer_idx = layer_idx
    return blk


def _init_weights(module, n_layer, initializer_range=0.02, rescale_prenorm_residual=True, n_residuals_per_layer=1):
    if isinstance(module, nn.Linear):
        if module.bias is not None and not getattr(module.bias, "_no_reinit", False):
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Embedding):
        nn.init.normal_(module.weight, std=initializer_range)
    if rescale_prenorm_residual:
        for name, p in module.named_parameters():
            if name in ("out_proj.weight", "fc2.weight"):
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                with torch.no_grad():
                    p /= math.sqrt(n_residuals_per_layer * n_layer)


@dataclass
class InferenceParams:
    max_seqlen: int
    max_batch_size: int
    seqlen_offset: int = 0
    batch_size_offset: int = 0
    key_value_memory_dict: dict = field(default_factory=dict)
    lengths_per_sample: Optional[torch.Tensor] = None
