"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package ``cleanumamba_b200``.

A self-contained, functional PyTorch-CPU restatement of the reference's hot path so that parity can be checked
on the GPU box (where /root/reference does not exist).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker / the CPU baseline.

What is restated (citations relative to /root/reference/):
  * ``CleanUMamba.valid_length / pad_signal``            src/network/CleanUMamba.py:219-246
  * ``CleanUMamba.forward`` (Mamba bottleneck branch)    src/network/CleanUMamba.py:252-324
  * GLU ``Activation.forward`` (bypass_channels == 0)    src/network/layers.py:26-41
  * streaming ``feed / _denoise_frame / flush``          src/network/CleanUMamba.py:358-490
  * the un-vendored dependency mamba-ssm==1.2.2 (environment.yml:29): ``Block.forward`` (non-fused),
    ``Mamba.forward`` slow path, ``Mamba.step`` and ``selective_scan_ref`` -- per SURVEY.md Appendix A.
The functions take a plain ``state_dict`` (reference key names) and derive every width from the tensor shapes,
so irregular pruned checkpoints need no special handling (cf. ``load_pruned_state_dict`` :492-550).

PINNING: this restatement is pinned against the unmodified reference imported from /root/reference on top of
``oracle/ref_shim`` (see ``oracle/make_golden.py`` -> ``tests/golden/*.pt`` and ``tests/test_oracle_pin.py``),
and the Mamba mixer additionally against ``transformers.models.mamba.modeling_mamba.MambaMixer.slow_forward``.
The reference itself ships no golden vectors, and the real mamba_ssm wheel is not installable
here; the pin is therefore "reference Python + restated dependency", stated as such in DESIGN.md.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------------------
# shape helpers
# --------------------------------------------------------------------------------------------------------------
def valid_length(length: int, depth: int, kernel: int = 4, stride: int = 2) -> int:
    """CleanUMamba.py:225-246 -- smallest length >= ``length`` that every conv / transposed conv maps exactly."""
    n = length
    for _ in range(depth):
        n = 1 if n < kernel else 1 + math.ceil((n - kernel) / stride)
    for _ in range(depth):
        n = (n - 1) * stride + kernel
    return int(n)


def model_dims(sd: StateDict) -> dict:
    """Everything the forward needs, read off the state_dict (cf. CleanUMamba.py:104-145, 540-545)."""
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("encoder."))
    n_mamba = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("tsfm_Mamba_layers."))
    kernel = sd["encoder.0.0.weight"].shape[2]
    layers = []
    for l in range(n_mamba):
        p = f"tsfm_Mamba_layers.{l}.mixer."
        d_inner, d_state = sd[p + "A_log"].shape
        layers.append(dict(d_inner=d_inner, d_state=d_state, dt_rank=sd[p + "dt_proj.weight"].shape[1],
                           d_conv=sd[p + "conv1d.weight"].shape[2]))
    return dict(depth=depth, kernel=kernel, n_mamba=n_mamba, d_model=sd["tsfm_conv1.weight"].shape[0],
                mamba=layers)


_GLU_ACT = {"Sigmoid": torch.sigmoid, "ReLU": torch.relu, "SiLU": F.silu, "GELU": F.gelu}


def glu(x: Tensor, activation: str = "Sigmoid") -> Tensor:
    """layers.py:26-34 with bypass_channels == 0: first half gated by act(second half)."""
    a, b = torch.split(x, [x.shape[1] // 2, x.shape[1] // 2], dim=1)
    return a * _GLU_ACT[activation](b)


def _cast(sd: StateDict, dtype, differentiable: bool = False) -> StateDict:
    if differentiable:      # gradient oracle: keep the autograd graph back to the caller's leaf tensors
        return {k: v.to(device="cpu", dtype=dtype) for k, v in sd.items()}
    return {k: v.detach().to(device="cpu", dtype=dtype) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------------------------
# Mamba block (mamba-ssm 1.2.2 slow path)
# --------------------------------------------------------------------------------------------------------------
def selective_scan(u: Tensor, delta: Tensor, A: Tensor, Bm: Tensor, Cm: Tensor, D: Optional[Tensor] = None,
                   z: Optional[Tensor] = None, delta_bias: Optional[Tensor] = None, delta_softplus: bool = False,
                   h0: Optional[Tensor] = None, return_last_state: bool = False):
    """``selective_scan_ref`` arithmetic: delta=softplus(delta+bias); h_t = exp(delta_t A) h_{t-1} + delta_t B_t u_t;
    y_t = <h_t, C_t> + D u_t; y_t *= silu(z_t).   u, delta, z: (b,d,l); A: (d,n); Bm, Cm: (b,n,l).
    Same recurrence as the reference loop, but the (b,d,l,n) tensors are never materialised (memory-lean port);
    ``h0`` (b,d,n) is an extension used by the streaming oracle (the reference always starts from zero)."""
    dt = u.dtype if u.dtype == torch.float64 else torch.float32
    out_dtype = u.dtype
    u, delta, A, Bm, Cm = u.to(dt), delta.to(dt), A.to(dt), Bm.to(dt), Cm.to(dt)
    if delta_bias is not None:
        delta = delta + delta_bias.to(dt)[..., None]
    if delta_softplus:
        delta = F.softplus(delta)
    b, d, l = u.shape
    h = torch.zeros(b, d, A.shape[1], dtype=dt) if h0 is None else h0.to(dt).clone()
    du = delta * u
    ys = []
    for t in range(l):
        h = torch.exp(delta[:, :, t, None] * A) * h + du[:, :, t, None] * Bm[:, None, :, t]
        ys.append(torch.einsum("bdn,bn->bd", h, Cm[:, :, t]))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u * D.to(dt)[..., None]
    if z is not None:
        y = y * F.silu(z.to(dt))
    y = y.to(out_dtype)
    return (y, h) if return_last_state else y


def mamba_mixer(h: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """``Mamba.forward`` with use_fast_path=False and no inference cache.  h: (b, l, d_model) -> same shape."""
    W = lambda name: sd[prefix + name]  # noqa: E731
    b, l, _ = h.shape
    d_inner, d_state = W("A_log").shape
    dt_rank = W("dt_proj.weight").shape[1]
    xz = (h @ W("in_proj.weight").t()).transpose(1, 2)                    # (b, 2*d_inner, l)
    x, z = xz[:, :d_inner], xz[:, d_inner:]
    k = W("conv1d.weight").shape[2]
    x = F.silu(F.conv1d(x, W("conv1d.weight"), W("conv1d.bias"), padding=k - 1, groups=d_inner)[..., :l])
    x_dbl = x.transpose(1, 2) @ W("x_proj.weight").t()                     # (b, l, R + 2N)
    dt_low, Bm, Cm = torch.split(x_dbl, [dt_rank, d_state, d_state], dim=-1)
    delta = (dt_low @ W("dt_proj.weight").t()).transpose(1, 2)            # bias is added inside the scan
    A = -torch.exp(W("A_log").float() if h.dtype != torch.float64 else W("A_log"))
    y = selective_scan(x, delta, A, Bm.transpose(1, 2), Cm.transpose(1, 2), W("D"), z=z,
                       delta_bias=W("dt_proj.bias"), delta_softplus=True)
    return y.transpose(1, 2) @ W("out_proj.weight").t()


def mamba_step(h: Tensor, conv_state: Tensor, ssm_state: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """``Mamba.step`` (einsum branch): one token.  h: (b, d_model); states updated in place."""
    W = lambda name: sd[prefix + name]  # noqa: E731
    d_inner, d_state = W("A_log").shape
    dt_rank = W("dt_proj.weight").shape[1]
    xz = h @ W("in_proj.weight").t()
    x, z = xz[:, :d_inner], xz[:, d_inner:]
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x
    x = F.silu((conv_state * W("conv1d.weight")[:, 0]).sum(-1) + W("conv1d.bias"))
    x_dbl = x @ W("x_proj.weight").t()
    dt_low, Bm, Cm = torch.split(x_dbl, [dt_rank, d_state, d_state], dim=-1)
    delta = F.softplus(dt_low @ W("dt_proj.weight").t() + W("dt_proj.bias"))
    A = -torch.exp(W("A_log"))
    ssm_state.copy_(ssm_state * torch.exp(delta[:, :, None] * A) + x[:, :, None] * (delta[:, :, None] * Bm[:, None, :]))
    y = torch.einsum("bdn,bn->bd", ssm_state, Cm) + W("D") * x
    return (y * F.silu(z)) @ W("out_proj.weight").t()


def _layer_norm(x: Tensor, sd: StateDict, prefix: str, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def bottleneck(x: Tensor, sd: StateDict, n_mamba: int, eps: float = 1e-5, step_states=None) -> Tensor:
    """tsfm_conv1 -> n x Block(pre-norm, fp32 residual) -> add + norm_f -> tsfm_conv2  (CleanUMamba.py:277-310).
    x: (b, C, l) -> (b, C, l).  ``step_states`` = list of (conv_state, ssm_state) selects the 1-token path."""
    x = F.conv1d(x, sd["tsfm_conv1.weight"], sd["tsfm_conv1.bias"])
    h = x.transpose(1, 2)
    residual = None
    for l in range(n_mamba):
        residual = h if residual is None else h + residual
        hn = _layer_norm(residual, sd, f"tsfm_Mamba_layers.{l}.norm.", eps)
        if step_states is None:
            h = mamba_mixer(hn, sd, f"tsfm_Mamba_layers.{l}.mixer.")
        else:
            assert hn.shape[1] == 1
            h = mamba_step(hn[:, 0], step_states[l][0], step_states[l][1], sd, f"tsfm_Mamba_layers.{l}.mixer.")[:, None]
    residual = h + residual
    h = _layer_norm(residual, sd, "norm_f.", eps)
    return F.conv1d(h.transpose(1, 2), sd["tsfm_conv2.weight"], sd["tsfm_conv2.bias"])


# --------------------------------------------------------------------------------------------------------------
# offline forward  (CleanUMamba.py:252-324)
# --------------------------------------------------------------------------------------------------------------
def forward(sd: StateDict, noisy: Tensor, *, stride: int = 2, normalize_input: bool = True, eps: float = 1e-5,
            glu_activation: str = "Sigmoid", dtype=torch.float32, return_intermediates: bool = False,
            differentiable: bool = False):
    """noisy: (B, L) or (B, 1, L) -> denoised (B, 1, L).  Does NOT mutate ``noisy`` (the reference does, :262).
    ``differentiable=True`` keeps the autograd graph to the tensors of ``sd`` (oracle for the backward kernels)."""
    sd = _cast(sd, dtype, differentiable)
    dims = model_dims(sd)
    D, K = dims["depth"], dims["kernel"]
    x = noisy.detach().to("cpu", dtype)
    if x.dim() == 2:
        x = x[:, None]
    assert x.shape[1] == 1
    L = x.shape[-1]
    std = None
    if normalize_input:
        std = x.std(dim=2, keepdim=True) + 1e-3
        x = x / std
    x = F.pad(x, (0, valid_length(L, D, K, stride) - L))
    skips: List[Tensor] = []
    for i in range(D):
        x = F.relu(F.conv1d(x, sd[f"encoder.{i}.0.weight"], sd[f"encoder.{i}.0.bias"], stride=stride))
        x = glu(F.conv1d(x, sd[f"encoder.{i}.2.weight"], sd[f"encoder.{i}.2.bias"]), glu_activation)
        skips.append(x)
    inter = {"skips": skips}
    x = bottleneck(x, sd, dims["n_mamba"], eps)
    inter["tsfm_out"] = x
    for j in range(D):
        skip = skips[D - 1 - j]
        x = x + skip[..., : x.shape[-1]]
        x = glu(F.conv1d(x, sd[f"decoder.{j}.0.weight"], sd[f"decoder.{j}.0.bias"]), glu_activation)
        x = F.conv_transpose1d(x, sd[f"decoder.{j}.2.weight"], sd[f"decoder.{j}.2.bias"], stride=stride)
        if j != D - 1:
            x = F.relu(x)
    x = x[..., :L] * std if normalize_input else x
    return (x, inter) if return_intermediates else x


# --------------------------------------------------------------------------------------------------------------
# streaming  (CleanUMamba.py:358-490)
# --------------------------------------------------------------------------------------------------------------
class StreamOracle:
    """Frame-by-frame restatement of ``feed`` / ``_denoise_frame`` / ``flush`` for a batch of independent streams.

    ``compat_skip_order_bug=True`` reproduces the shipped indexing ``skip_connections[i]`` (:474, not reversed) so
    the restatement can be pinned against the unmodified reference on equal-width models; the default ``False`` uses
    the intended ``skip_connections[-1-i]`` for which streaming == offline ``forward`` on every emitted sample."""

    def __init__(self, sd: StateDict, *, batch: int = 1, stride: int = 2, normalize_input: bool = True,
                 eps: float = 1e-5, glu_activation: str = "Sigmoid", dtype=torch.float32,
                 compat_skip_order_bug: bool = False):
        self.sd = _cast(sd, dtype)
        self.dims = model_dims(self.sd)
        self.D, self.K, self.S = self.dims["depth"], self.dims["kernel"], stride
        self.batch, self.dtype, self.eps, self.act = batch, dtype, eps, glu_activation
        self.normalize_input, self.bug = normalize_input, compat_skip_order_bug
        self.frame_length = valid_length(1, self.D, self.K, self.S)
        self.total_stride = self.S ** self.D
        self.frames = 0
        self.input_std = torch.zeros(batch, 1, dtype=dtype)
        self.pending = torch.zeros(batch, 0, dtype=dtype)
        self.cache: Dict[str, Tensor] = {}
        d_model = self.dims["d_model"]
        self.states = [(torch.zeros(batch, m["d_inner"], m["d_conv"], dtype=dtype),
                        torch.zeros(batch, m["d_inner"], m["d_state"], dtype=dtype)) for m in self.dims["mamba"]]
        del d_model

    def _enc(self, i: int, x: Tensor) -> Tensor:
        sd = self.sd
        x = F.relu(F.conv1d(x, sd[f"encoder.{i}.0.weight"], sd[f"encoder.{i}.0.bias"], stride=self.S))
        return glu(F.conv1d(x, sd[f"encoder.{i}.2.weight"], sd[f"encoder.{i}.2.bias"]), self.act)

    def _frame(self, frame: Tensor) -> Tensor:
        sd, D, K, S = self.sd, self.D, self.K, self.S
        x = frame[:, None]
        skips = []
        hop = self.total_stride
        for i in range(D):
            hop //= S
            length = x.shape[2]
            prev = self.cache.get(f"enc{i}")
            if prev is not None:       # only the not-yet-computed output columns (:432-440)
                x = x[..., length - K - S * ((length - K) // S - prev.shape[-1]):]
            x = self._enc(i, x)
            if prev is not None:
                x = torch.cat([prev, x], -1)
            self.cache[f"enc{i}"] = x[..., hop:]
            skips.append(x)
        x = bottleneck(x, sd, self.dims["n_mamba"], self.eps, step_states=self.states)
        for j in range(D):
            skip = skips[j] if self.bug else skips[D - 1 - j]
            x = x + skip[..., : x.shape[-1]]
            x = glu(F.conv1d(x, sd[f"decoder.{j}.0.weight"], sd[f"decoder.{j}.0.bias"]), self.act)
            bias = sd[f"decoder.{j}.2.bias"]
            x = F.conv_transpose1d(x, sd[f"decoder.{j}.2.weight"], bias, stride=S)
            prev = self.cache.get(f"dec{j}")
            self.cache[f"dec{j}"] = x[..., -S:] - bias.view(-1, 1)          # overlap-add tail, bias removed (:480)
            x = x[..., :-S].clone()
            if prev is not None:
                x[..., :S] += prev
            if j != D - 1:
                x = F.relu(x)
        return x[:, 0]

    def feed(self, chunk: Tensor) -> Tensor:
        """chunk: (batch, n) -> (batch, m) with m = total_stride * (#complete frames now available)."""
        self.pending = torch.cat([self.pending, chunk.detach().to("cpu", self.dtype)], dim=1)
        outs = []
        while self.pending.shape[1] >= self.frame_length:
            self.frames += 1
            frame = self.pending[:, : self.frame_length]
            if self.normalize_input:          # running arithmetic mean of the per-frame std (:399-401)
                self.input_std = (frame.std(dim=1, keepdim=True) + 1e-3) / self.frames \
                    + (1 - 1 / self.frames) * self.input_std
                frame = frame / self.input_std
            out = self._frame(frame)[:, : self.total_stride]
            if self.normalize_input:
                out = out * self.input_std
            outs.append(out)
            self.pending = self.pending[:, self.total_stride:]
        return torch.cat(outs, 1) if outs else torch.zeros(self.batch, 0, dtype=self.dtype)

    def flush(self) -> Tensor:
        """:358-368 -- clears the conv caches (not the Mamba state), feeds frame_length zeros, returns the tail."""
        self.cache = {}
        n = self.pending.shape[1]
        return self.feed(torch.zeros(self.batch, self.frame_length, dtype=self.dtype))[:, :n]


# --------------------------------------------------------------------------------------------------------------
# pruning: channel importances  (src/pruning/pruninggroup.py:160-226)
# --------------------------------------------------------------------------------------------------------------
def channel_importances(weight: Tensor, grad: Optional[Tensor], dim: int = 0, channel_offset: int = 0,
                        n_channels: Optional[int] = None, n_heads: int = 1) -> dict:
    """Restatement of ``PruningModule.channel_importances`` (pruninggroup.py:160-226) without the module bookkeeping:
    transpose for dim 1 (:178-181), flatten the trailing dims (:189-196), cut the group's rows out at ``channel_offset``
    (:199-204), fold ``n_heads`` rows into one channel, then the five sums (:213-221)."""
    w = weight.detach().to("cpu", torch.float32)
    g = None if grad is None else grad.detach().to("cpu", torch.float32)
    if dim == 1:
        w = w.transpose(1, 0)
        g = None if g is None else g.transpose(1, 0)
    if w.dim() > 2:
        w = w.flatten(1)
        g = None if g is None else g.flatten(1)
    elif w.dim() == 1:
        w = w.unsqueeze(1)
        g = None if g is None else g.unsqueeze(1)
    if n_channels is None:
        n_channels = (w.shape[0] - channel_offset) // n_heads
    rows = slice(channel_offset, channel_offset + n_channels * n_heads)
    w = w[rows].reshape(n_channels, -1)
    out = {"weight": w.abs().pow(2).sum(1), "grad": None, "taylor_individual": None, "taylor_squared_individual": None,
           "taylor_group": None, "n_parameters": w.shape[1], "act_var": None}
    if g is not None:
        g = g[rows].reshape(n_channels, -1)
        out["grad"] = g.abs().pow(2).sum(1)
        out["taylor_individual"] = (w * g).abs().sum(1)
        out["taylor_squared_individual"] = (w * g).pow(2).sum(1)
        out["taylor_group"] = (w * g).sum(1).abs()
    return out


# --------------------------------------------------------------------------------------------------------------
# metrics used by the parity tests and the bench
# --------------------------------------------------------------------------------------------------------------
def si_sdr(estimate: Tensor, target: Tensor) -> Tensor:
    """Scale-invariant SDR in dB per clip; inputs (B, L) or (B, 1, L)."""
    e = estimate.reshape(estimate.shape[0], -1).double()
    t = target.reshape(target.shape[0], -1).double()
    e = e - e.mean(-1, keepdim=True)
    t = t - t.mean(-1, keepdim=True)
    proj = (e * t).sum(-1, keepdim=True) / (t.pow(2).sum(-1, keepdim=True) + 1e-20) * t
    return 10 * torch.log10(proj.pow(2).sum(-1) / ((e - proj).pow(2).sum(-1) + 1e-20) + 1e-20)


def synth_batch(batch: int, seconds: float, seed: int = 1234, sr: int = 16000) -> Tuple[Tensor, Tensor]:
    """Synthetic (clean, noisy) speech-like mixtures of SURVEY.md §8(d): 8 harmonics of f0~U(80,300) Hz with a 4 Hz
    raised-cosine envelope, peak 0.3, plus 1-pole low-passed white noise at SNR~U(-5,25) dB.  (B,1,T) fp32."""
    g = torch.Generator().manual_seed(seed)
    T = int(round(seconds * sr))
    t = torch.arange(T, dtype=torch.float64) / sr
    f0 = 80 + 220 * torch.rand(batch, 1, generator=g, dtype=torch.float64)
    phase = 2 * math.pi * torch.rand(batch, 8, generator=g, dtype=torch.float64)
    amp = 1.0 / torch.arange(1, 9, dtype=torch.float64)
    clean = torch.zeros(batch, T, dtype=torch.float64)
    for k in range(8):
        clean += amp[k] * torch.sin(2 * math.pi * (k + 1) * f0 * t + phase[:, k:k + 1])
    env_phase = 2 * math.pi * torch.rand(batch, 1, generator=g, dtype=torch.float64)
    clean *= 0.5 * (1 - torch.cos(2 * math.pi * 4.0 * t + env_phase))
    clean *= 0.3 / clean.abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
    white = torch.randn(batch, T, generator=g, dtype=torch.float64)
    alpha = 0.95 * torch.rand(batch, generator=g, dtype=torch.float64)
    # 1-pole low-pass y[t] = a*y[t-1] + (1-a)*x[t], evaluated as a truncated FIR (a^64 < 4e-2 at a=0.95; 256 taps)
    taps = 256
    kern = (1 - alpha)[:, None] * alpha[:, None] ** torch.arange(taps, dtype=torch.float64)[None]
    noise = F.conv1d(F.pad(white[None], (taps - 1, 0)), kern.flip(1)[:, None], groups=batch)[0]
    snr_db = -5 + 30 * torch.rand(batch, 1, generator=g, dtype=torch.float64)
    scale = clean.pow(2).mean(1, keepdim=True).sqrt() / noise.pow(2).mean(1, keepdim=True).sqrt().clamp_min(1e-12)
    noise = noise * scale * 10 ** (-snr_db / 20)
    return clean.float()[:, None], (clean + noise).float()[:, None]
