"""Probe: two (four) part-batch time-major sessions on two (four) CUDA streams inside ONE graph (HBM-bound state updates of one part
under the tensor-bound GEMMs of another) vs one full-batch session.  E6 full, S streams in total, H hops per call.
Measured on B200 (4096 streams): 1 hop 3.98 ms (one session) vs 4.51 (2 parts) / 4.14 (2 parts, second part started after the first
part's encoder via an event inside _process) / 4.65 (4 parts); 2 hops 5.27 vs 5.64 / 5.49 / 5.84 -- SLOWER: a GEMM CTA needs a whole
SM's shared memory and cannot co-reside with the state-update CTAs, and the part-batch GEMMs fill the machine worse.  Not adopted."""
import sys

import torch

sys.path.insert(0, ".")
from bench import CONFIGS  # noqa: E402
from cleanumamba_b200.network import Net  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Net("CleanUMamba", dict(CONFIGS["e6"], math_mode="f16x3")).to(dev).eval()
hop, fl = net.total_stride, net.frame_length
n = H * hop


def prime(sess, B):
    sess.feed(torch.randn(B, fl - hop, device=dev) * 0.1)
    c = torch.randn(B, n, device=dev) * 0.1
    for _ in range(4):
        sess.feed(c)
    return c


def timeit(fn, reps=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- reference: one session, graph
one = net.stream_session(batch=S, layout="time_major")
c = prime(one, S)
one.capture_graph(n)
t_one = timeit(lambda: one.feed(c))
print(f"one session of {S}: {t_one:.3f} ms per call", flush=True)
del one
torch.cuda.empty_cache()

for parts, dephase in ((2, False), (4, False)):
    B = S // parts
    sess = [net.stream_session(batch=B, layout="time_major") for _ in range(parts)]
    chunks = [prime(s_, B) for s_ in sess]
    side = [torch.cuda.Stream() for _ in range(parts - 1)]
    F = n // hop
    # static inputs already inside each session's x_buf: copy chunk, then replay
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    outs = [None] * parts
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        evs = [torch.cuda.Event() for _ in range(parts)]
        for k in range(parts):
            st = cur if k == 0 else side[k - 1]
            if k > 0:
                if dephase:
                    st.wait_event(evs[k - 1])          # start once the previous part's encoder is queued/done
                else:
                    st.wait_stream(cur) if k == 1 else st.wait_stream(cur)
            with torch.cuda.stream(st):
                s_ = sess[k]
                if dephase:
                    s_.after_encoder = (lambda e=evs[k], stream=st: e.record(stream))
                if k == 0 and not dephase:
                    pass
                n_total = s_.n_pend + n
                outs[k] = s_._process(F, n_total)
                s_.after_encoder = None
        for sd in side:
            cur.wait_stream(sd)

    def step():
        for k, s_ in enumerate(sess):
            s_.x_buf[:, s_.n_pend: s_.n_pend + n].copy_(chunks[k])
        g.replay()
        # host counters are steady-state invariant for timing purposes (the graph has fixed addresses); skip _advance

    t = timeit(step)
    print(f"{parts} sessions of {B} on {parts} streams, dephase={dephase}: {t:.3f} ms per call ({n / 16.0 / t:.2f}x real time)", flush=True)
    del sess, g
    torch.cuda.empty_cache()
