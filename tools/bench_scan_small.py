"""Small-batch selective scan: time-sequential vs segment-parallel (E8 geometry: d_inner 2048, d_state 64), CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cleanumamba_b200 import ops
dev = "cuda"
for b, sec in [(1, 10), (1, 30), (1, 60), (4, 10), (2, 60), (8, 10)]:
    l = {10: 624, 30: 1874, 60: 3749}[sec]
    d, n = 2048, 64
    g = torch.Generator(device=dev).manual_seed(0)
    mk = lambda *s: torch.randn(*s, generator=g, device=dev)
    u, delta, Bm, Cm, z = mk(b, d, l), mk(b, d, l) * 0.5, mk(b, n, l), mk(b, n, l), mk(b, d, l)
    A, D, bias = -torch.exp(mk(d, n) * 0.5 + 0.5), mk(d), mk(d) * 0.5 - 2
    out = {}
    for seg in (False, True):
        for _ in range(3):
            y = ops.selective_scan_fn(u, delta, A, Bm, Cm, D, z, bias, True, segment_parallel=seg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = ops.selective_scan_fn(u, delta, A, Bm, Cm, D, z, bias, True, segment_parallel=seg)
        e1.record(); torch.cuda.synchronize()
        out[seg] = (e0.elapsed_time(e1) / 10, y)
    print(f"batch {b} x {sec:2d} s (L={l}): sequential {out[False][0]:.3f} ms  segment-parallel {out[True][0]:.3f} ms  "
          f"(incl. the (b,d,l)->(b,l,d) layout transposes of the operator-level wrapper)  max|dy| {(out[False][1]-out[True][1]).abs().max().item():.2e}")
