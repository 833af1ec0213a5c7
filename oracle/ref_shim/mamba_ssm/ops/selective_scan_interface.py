"""Oracle shim: PyTorch restatement of mamba_ssm.ops.selective_scan_interface.selective_scan_ref
(mamba-ssm 1.2.2; semantics per SURVEY.md Appendix A).  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False):
    """u, delta: (b, d, l); A: (d, n); B, C: (b, n, l) [input-dependent, real]; D: (d,); z: (b, d, l)."""
    in_dtype = u.dtype
    u = u.float()
    delta = delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    b, d, l = u.shape
    n = A.shape[1]
    B = B.float()
    C = C.float()
    decay = torch.exp(torch.einsum("bdl,dn->bdln", delta, A))
    drive = torch.einsum("bdl,bnl,bdl->bdln", delta, B, u)
    h = A.new_zeros((b, d, n))
    ys = []
    for t in range(l):
        h = decay[:, :, t] * h + drive[:, :, t]
        ys.append(torch.einsum("bdn,bn->bd", h, C[:, :, t]))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D[..., None]
    if z is not None:
        out = out * F.silu(z)
    out = out.to(dtype=in_dtype)
    return out if not return_last_state else (out, h)


# the reference's CPU path: the CUDA op is replaced by the PyTorch one
selective_scan_fn = selective_scan_ref
mamba_inner_fn = None
