"""First-contact probe of the CTA-pair GEMM on a GPU box: correctness vs the single-CTA kernel, then timing of the
E8 layer shapes with both.  Run under `timeout`; every wait in the kernel is bounded (traps instead of hanging)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from cleanumamba_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(math, b, k, n, m, taps, epi, cta_pair, iters=0):
    rows = m + 1 if taps == 2 else m
    a = torch.randn(b, rows, k, device=dev)
    w = torch.randn(taps, n, k, device=dev) / (taps * k) ** 0.5
    bias = torch.randn(n, device=dev)
    shifts = (0, 1) if taps == 2 else (0, 0)
    out = ops.gemm_bias_act(a, w, bias, epi, shifts=shifts, m=m, math=math, cta_pair=cta_pair)
    torch.cuda.synchronize()
    return out, (a, w, bias, shifts)


if len(sys.argv) > 2 and sys.argv[1] == "--one":
    # for ncu: the e4c shape once on a CTA pair, once on single CTAs
    b, k, n, m, taps = 64, 1024, 768, 5006, 2
    for cp in (1, -1):
        a = torch.randn(b, m + 1, k, device=dev)
        w = torch.randn(taps, n, k, device=dev) / (2 * k) ** 0.5
        ops.gemm_bias_act(a, w, torch.randn(n, device=dev), _lib.EPI_RELU, shifts=(0, 1), m=m, math=sys.argv[2], cta_pair=cp)
        torch.cuda.synchronize()
    sys.exit(0)

print("== correctness", flush=True)
for math in ("f16x3", "tf32x3", "tf32", "bf16x3"):
    for (b, k, n, m, taps) in [(1, 64, 256, 129, 1), (1, 64, 256, 256, 1), (2, 128, 768, 700, 1), (1, 96, 160, 300, 1),
                               (3, 256, 512, 257, 2), (1, 512, 1536, 5000, 1)]:
        torch.manual_seed(1)
        p, _ = run(math, b, k, n, m, taps, _lib.EPI_RELU, 1)
        torch.manual_seed(1)
        s, _ = run(math, b, k, n, m, taps, _lib.EPI_RELU, -1)
        print(math, (b, k, n, m, taps), "max|pair-single| =", (p - s).abs().max().item(), flush=True)

print("== timing (ms): E8 layer shapes, batch 64 x 10 s", flush=True)
shapes = [("e2c", 64, 512, 256, 20030, 2), ("e2g", 64, 256, 512, 20030, 1), ("e3c", 64, 1024, 512, 10014, 2),
          ("e3g", 64, 512, 1024, 10014, 1), ("e4c", 64, 2048, 768, 5006, 2), ("e4g", 64, 768, 1536, 5006, 1),
          ("e5c", 64, 3072, 768, 2502, 2), ("m_in", 1, 512, 4096, 39936, 1)]
for math in ("f16x3",):
    for name, b, k, n, m, taps in shapes:
        res = {}
        for cp in (1, -1):
            rows = m + 1 if taps == 2 else m
            a = torch.randn(b, rows, k // taps, device=dev)
            w = torch.randn(taps, n, k // taps, device=dev) / k ** 0.5
            bias = torch.randn(n, device=dev)
            shifts = (0, 1) if taps == 2 else (0, 0)
            for _ in range(2):
                ops.gemm_bias_act(a, w, bias, _lib.EPI_RELU, shifts=shifts, m=m, math=math, cta_pair=cp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.gemm_bias_act(a, w, bias, _lib.EPI_RELU, shifts=shifts, m=m, math=math, cta_pair=cp)
            e1.record()
            torch.cuda.synchronize()
            res[cp] = e0.elapsed_time(e1) / 5
            del a, w
        print(f"{math} {name}: pair {res[1]:.3f}  single {res[-1]:.3f}  (includes the weight split + output alloc)", flush=True)
