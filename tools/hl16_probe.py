"""Timing of the f16x3 GEMM with fp32 vs pre-split (hl16) activations on E8 layer shapes (batch 64 x 10 s)."""
import sys

import torch

sys.path.insert(0, ".")
from cleanumamba_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [("e1g", 64, 128, 256, 40062, 1, "glu"), ("e2c", 64, 512, 256, 20030, 2, "relu"), ("e2g", 64, 256, 512, 20030, 1, "glu"),
          ("e3c", 64, 1024, 512, 10014, 2, "relu"), ("e4c", 64, 2048, 768, 5006, 2, "relu"), ("e4g", 64, 768, 1536, 5006, 1, "glu"),
          ("e0g", 64, 64, 128, 80126, 1, "glu")]
for name, b, k, n, m, taps, ep in shapes:
    epi = _lib.EPI_RELU if ep == "relu" else _lib.EPI_GLU["Sigmoid"]
    rows = m + 1 if taps == 2 else m
    a = torch.randn(b, rows, k // taps, device=dev)
    ah = ops.split_hl16(a)
    w = torch.randn(taps, n, k // taps, device=dev) / k ** 0.5
    bias = torch.randn(n, device=dev)
    shifts = (0, 1) if taps == 2 else (0, 0)
    res = {}
    for label, aa, oh in (("fp32->fp32", a, False), ("hl16->fp32", ah, False), ("hl16->hl16", ah, True), ("fp32->hl16", a, True)):
        for _ in range(2):
            ops.gemm_bias_act(aa, w, bias, epi, shifts=shifts, m=m, math="f16x3", out_hl16=oh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gemm_bias_act(aa, w, bias, epi, shifts=shifts, m=m, math="f16x3", out_hl16=oh)
        e1.record()
        torch.cuda.synchronize()
        res[label] = e0.elapsed_time(e1) / 5
    print(name, {k_: round(v, 3) for k_, v in res.items()}, flush=True)
    del a, ah, w
