"""Oracle shim: PyTorch restatement of mamba_ssm.modules.mamba_simple.{Mamba, Block}
(mamba-ssm 1.2.2, slow path only).  TEST INFRASTRUCTURE ONLY."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

causal_conv1d_fn = None
causal_conv1d_update = None
selective_state_update = None


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None):
        kw = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx

        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=bias, **kw)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, kernel_size=d_conv, groups=self.d_inner,
                                padding=d_conv - 1, bias=conv_bias, **kw)
        self.activation = "silu"
        self.act = nn.SiLU()
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False, **kw)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **kw)

        scale = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, scale)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -scale, scale)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(self.d_inner, **kw) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        inv_softplus_dt = dt + torch.log(-torch.expm1(-dt))
        with torch.no_grad():
            self.dt_proj.bias.copy_(inv_softplus_dt)
        self.dt_proj.bias._no_reinit = True

        A = torch.arange(1, d_state + 1, dtype=torch.float32, device=device)[None, :].expand(self.d_inner, -1)
        self.A_log = nn.Parameter(torch.log(A.contiguous()))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.d_inner, device=device))
        self.D._no_weight_decay = True
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **kw)

    def forward(self, hidden_states, inference_params=None):
        b, l, _ = hidden_states.shape
        conv_state = ssm_state = None
        if inference_params is not None:
            conv_state, ssm_state = self._get_states_from_cache(inference_params, b)
            if inference_params.seqlen_offset > 0:
                out, _, _ = self.step(hidden_states, conv_state, ssm_state)
                return out
        xz = (self.in_proj.weight @ hidden_states.reshape(b * l, -1).t()).reshape(-1, b, l).permute(1, 0, 2)
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(xz.dtype)[:, None]
        A = -torch.exp(self.A_log.float())
        x, z = xz.chunk(2, dim=1)
        if conv_state is not None:
            conv_state.copy_(F.pad(x, (self.d_conv - x.shape[-1], 0)))
        x = self.act(self.conv1d(x)[..., :l])
        x_dbl = self.x_proj(x.transpose(1, 2).reshape(b * l, -1))
        dt, Bm, Cm = torch.split(x_dbl, [self.dt_rank, self.d_state, self.d_state], dim=-1)
        dt = (self.dt_proj.weight @ dt.t()).reshape(-1, b, l).permute(1, 0, 2)
        Bm = Bm.reshape(b, l, -1).transpose(1, 2).contiguous()
        Cm = Cm.reshape(b, l, -1).transpose(1, 2).contiguous()
        y = selective_scan_fn(x, dt, A, Bm, Cm, self.D.float(), z=z, delta_bias=self.dt_proj.bias.float(),
                              delta_softplus=True, return_last_state=ssm_state is not None)
        if ssm_state is not None:
            y, last = y
            ssm_state.copy_(last)
        return self.out_proj(y.transpose(1, 2))

    def step(self, hidden_states, conv_state, ssm_state):
        dtype = hidden_states.dtype
        assert hidden_states.shape[1] == 1, "Only support decoding with 1 token at a time for now"
        xz = self.in_proj(hidden_states.squeeze(1))
        x, z = xz.chunk(2, dim=-1)
        conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
        conv_state[:, :, -1] = x
        x = torch.sum(conv_state * self.conv1d.weight.squeeze(1), dim=-1)
        if self.conv1d.bias is not None:
            x = x + self.conv1d.bias
        x = self.act(x).to(dtype=dtype)
        x_db = self.x_proj(x)
        dt, Bm, Cm = torch.split(x_db, [self.dt_rank, self.d_state, self.d_state], dim=-1)
        dt = F.linear(dt, self.dt_proj.weight)
        A = -torch.exp(self.A_log.float())
        dt = F.softplus(dt + self.dt_proj.bias.to(dtype=dt.dtype))
        dA = torch.exp(torch.einsum("bd,dn->bdn", dt, A))
        dB = torch.einsum("bd,bn->bdn", dt, Bm)
        ssm_state.copy_(ssm_state * dA + x[:, :, None] * dB)
        y = torch.einsum("bdn,bn->bd", ssm_state.to(dtype), Cm)
        y = y + self.D.to(dtype) * x
        y = y * self.act(z)
        return self.out_proj(y).unsqueeze(1), conv_state, ssm_state

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        dev = self.out_proj.weight.device
        conv = torch.zeros(batch_size, self.d_model * self.expand, self.d_conv, device=dev,
                           dtype=self.conv1d.weight.dtype if dtype is None else dtype)
        ssm = torch.zeros(batch_size, self.d_model * self.expand, self.d_state, device=dev,
                          dtype=self.dt_proj.weight.dtype if dtype is None else dtype)
        return conv, ssm

    def _get_states_from_cache(self, inference_params, batch_size, initialize_states=False):
        assert self.layer_idx is not None
        if self.layer_idx not in inference_params.key_value_memory_dict:
            inference_params.key_value_memory_dict[self.layer_idx] = self.allocate_inference_cache(batch_size, 1)
        conv_state, ssm_state = inference_params.key_value_memory_dict[self.layer_idx]
        if initialize_states:
            conv_state.zero_()
            ssm_state.zero_()
        return conv_state, ssm_state


class Block(nn.Module):
    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.mixer = mixer_cls(dim)
        self.norm = norm_cls(dim)
        assert not fused_add_norm, "oracle shim restates the non-fused (PyTorch) branch only"

    def forward(self, hidden_states, residual=None, inference_params=None):
        residual = (hidden_states + residual) if residual is not None else hidden_states
        hidden_states = self.norm(residual.to(dtype=self.norm.weight.dtype))
        if self.residual_in_fp32:
            residual = residual.to(torch.float32)
        return self.mixer(hidden_states, inference_params=inference_params), residual

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)
