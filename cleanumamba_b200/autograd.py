"""Training-mode forward: one torch.autograd.Function around the whole CUDA path.

The loss (L1 + multi-resolution STFT, /root/reference/src/util/util.py:215-327) stays in PyTorch and back-propagates
into ``out``; ``backward`` then runs our kernels (train_engine.TrainEngine.backward) and hands PyTorch one gradient per
parameter, so optimisers / GradScaler / gradient clipping / DDP-style all-reduce work unchanged.  No PyTorch autograd
fallback exists for the model body."""
import torch


class _CleanUMambaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, noisy, *params):
        eng = model.train_engine()
        with torch.no_grad():
            out, saved = eng.forward_train(noisy)
        ctx.eng, ctx.saved = eng, saved
        ctx.names = [n for n, _ in model.named_parameters()]
        ctx.needs = [p.requires_grad for p in params]
        ctx.dtypes = [p.dtype for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        with torch.no_grad():
            sync = getattr(ctx.eng.model, "_grad_sync", None)
            ctx.eng.backward(ctx.saved, dout, sync=sync)
            if sync is not None:
                scale = sync.finish()           # waits for the three bucket all-reduces (stream-ordered)
                if scale != 1.0:
                    ctx.eng.gflat.mul_(scale)
            grads = ctx.eng.unpack_grads()
        ctx.saved = None
        lo = ctx.eng.gflat.data_ptr()
        hi = lo + ctx.eng.gflat.numel() * 4

        def own(g, dt):       # a gradient handed to autograd must never alias the persistent buffer (it may become param.grad)
            g = g.to(dt)
            return g.clone() if lo <= g.data_ptr() < hi else g
        outs = [own(grads[n], dt) if need else None for n, need, dt in zip(ctx.names, ctx.needs, ctx.dtypes)]
        return (None, None, *outs)


def forward_with_grad(model, noisy_audio, return_skip_connections=False):
    if return_skip_connections:
        raise NotImplementedError("cleanumamba_b200: return_skip_connections is inference-only (knowledge-distillation "
                                  "losses of the reference are outside the hot path)")
    return _CleanUMambaFn.apply(model, noisy_audio, *list(model.parameters()))
