"""GPU parity of the carried-state streaming path (StreamSession / feed / flush) against the CPU streaming oracle
(oracle.StreamOracle == the reference's feed/_denoise_frame/flush with the skip order fixed) and against offline forward."""
import json

import pytest
import torch

import cleanumamba_oracle as orc
from conftest import load_golden

pytestmark = pytest.mark.gpu
TOL = 2e-5      # fp32, outputs are O(0.1); chunked vs frame-by-frame differ only by summation order


@pytest.fixture(autouse=True, params=["time_major", "stream_major"])
def default_session_layout(request):
    """Every test of this file runs on both carried-state sessions: stream_session(layout="auto") -- and with it the module-level
    feed() / flush() -- picks the time-major session from TIME_MAJOR_MIN_STREAMS streams (1: always) or the stream-major one."""
    from cleanumamba_b200.CleanUMamba import CleanUMamba
    old = CleanUMamba.TIME_MAJOR_MIN_STREAMS
    CleanUMamba.TIME_MAJOR_MIN_STREAMS = 1 if request.param == "time_major" else 10 ** 9
    yield request.param
    CleanUMamba.TIME_MAJOR_MIN_STREAMS = old


def build(fx, **kw):
    from cleanumamba_b200.network import Net
    net = Net("CleanUMamba", {**json.loads(fx["config"]), **kw})
    net.load_pruned_state_dict(fx["state_dict"])
    return net.cuda().float().eval()


def ragged_chunks(total, seed, lo=1, hi=1500):
    g = torch.Generator().manual_seed(seed)
    cuts, pos = [], 0
    while pos < total:
        n = int(torch.randint(lo, hi, (1,), generator=g))
        cuts.append((pos, min(total, pos + n)))
        pos += n
    return cuts


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("name,math", [("tiny_equalwidth_seed0", "fp32"), ("e6_pruned_200k", "fp32"),
                                       ("e8_pruned_500k", "fp32"), ("e8_pruned_500k", "tf32x3"), ("e8_pruned_500k", "bf16x3"), ("e8_pruned_500k", "f16x3")])
def test_stream_matches_stream_oracle(name, math, normalize):
    fx = load_golden(name)
    net = build(fx, normalize_input=normalize, math_mode=math)
    x = fx["noisy"][:1, 0]                                  # (1, T)
    x = x[:, : min(x.shape[1], 6000)]
    so = orc.StreamOracle(fx["state_dict"], normalize_input=normalize)
    want = torch.cat([so.feed(x), so.flush()], 1)
    sess = net.stream_session(batch=1)
    outs = [sess.feed(x[:, a:b].cuda()) for a, b in ragged_chunks(x.shape[1], 3)]
    outs.append(sess.flush())
    got = torch.cat(outs, 1).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < TOL
    if not normalize:       # streaming == offline forward on every sample emitted before the flush
        off = orc.forward(fx["state_dict"], x, normalize_input=False)[:, 0]
        n_emit = sum(o.shape[1] for o in outs[:-1])
        assert n_emit > 0 and (got[:, :n_emit] - off[:, :n_emit]).abs().max().item() < TOL


def test_batched_streams_multi_hop_chunks():
    """4 independent streams, 8 hops per call: equals 4 single-stream oracles."""
    fx = load_golden("e6_pruned_200k")
    net = build(fx, normalize_input=True)
    g = torch.Generator().manual_seed(11)
    B, hop = 4, 64
    x = torch.randn(B, 190 + hop * 8 * 5, generator=g) * 0.1 * (1 + torch.arange(B)[:, None])
    sess = net.stream_session(batch=B)
    outs, pos = [], 0
    for n in [190 + hop * 7] + [hop * 8] * 4 + [hop * 1]:
        outs.append(sess.feed(x[:, pos:pos + n].cuda()))
        pos += n
    got = torch.cat(outs, 1).cpu()
    for b in range(B):
        so = orc.StreamOracle(fx["state_dict"], normalize_input=True)
        want = so.feed(x[b:b + 1, :pos])
        assert got[b:b + 1].shape == want.shape
        assert (got[b:b + 1] - want).abs().max().item() < TOL * (1 + b)


def test_batched_streams_large_and_small_chunks_mix_gemm_modes():
    """3 streams fed 40-hop chunks (>= 128 new rows per stream at the outer levels: one GEMM batch item per stream, operands read
    and written in place inside the FIFOs), then 1- and 2-hop chunks (flattened streams, state-update scan kernel), then large
    again: every transition between the two operand layouts, against offline forward on the fed signal."""
    fx = load_golden("e6_pruned_200k")
    net = build(fx, normalize_input=False, math_mode="f16x3")
    g = torch.Generator().manual_seed(5)
    B, hop = 3, 64
    sizes = [190 + hop * 39, hop * 40, hop * 1, hop * 2, hop * 1, hop * 40, hop * 3, hop * 40]
    x = torch.randn(B, sum(sizes), generator=g) * 0.1
    sess = net.stream_session(batch=B)
    outs, pos = [], 0
    for n in sizes:
        outs.append(sess.feed(x[:, pos:pos + n].cuda()))
        pos += n
    got = torch.cat(outs, 1).cpu()
    off = orc.forward(fx["state_dict"], x, normalize_input=False)[:, 0]
    assert got.shape[1] > 0.9 * x.shape[1]
    assert (got - off[:, : got.shape[1]]).abs().max().item() < 1e-4      # f16x3 products (BASELINE tolerance)


def test_single_hop_feeds_use_state_update_scan_d_state_64():
    """d_state = 64 (the shipped full-size geometry): 1- and 2-hop feeds run the state-update scan kernel; streamed output
    equals offline forward of the same weights (CPU oracle)."""
    from cleanumamba_b200.network import Net
    torch.manual_seed(3)
    cfg = dict(channels_input=1, channels_output=1, channels_H=16, max_H=32, encoder_n_layers=3, kernel_size=4, stride=2,
               tsfm_n_layers=2, tsfm_n_head=1, tsfm_d_model=64, tsfm_d_inner=64, normalize_input=False, math_mode="fp32")
    net = Net("CleanUMamba", cfg).cuda().float().eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    B, hop = 2, 8
    g = torch.Generator().manual_seed(9)
    sizes = [net.frame_length] + [hop] * 9 + [2 * hop] * 4 + [hop * 5] + [hop] * 3
    x = torch.randn(B, sum(sizes), generator=g) * 0.1
    sess = net.stream_session(batch=B)
    outs, pos = [], 0
    for n in sizes:
        outs.append(sess.feed(x[:, pos:pos + n].cuda()))
        pos += n
    got = torch.cat(outs, 1).cpu()
    off = orc.forward(sd, x, normalize_input=False)[:, 0]
    assert got.shape[1] >= x.shape[1] - net.frame_length
    assert (got - off[:, : got.shape[1]]).abs().max().item() < TOL


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("batch,hops", [(1, 1), (3, 4), (2, 40)])
def test_cuda_graph_replay_equals_eager_feed(batch, hops, normalize):
    """capture_graph(): the steady-state feed() replayed from a CUDA graph (static state tensors, device-side frame counter)
    gives bit-identical output to the eager path over many steps, survives a chunk-size change (falls back to eager, state
    intact) and can be re-captured."""
    fx = load_golden("e6_pruned_200k")
    hop = 64
    g = torch.Generator().manual_seed(21)
    n = hops * hop
    x = torch.randn(batch, 190 - hop + n * 9 + 3 * hop, generator=g) * 0.1
    outs = {}
    for mode in ("eager", "graph"):
        net = build(fx, normalize_input=normalize, math_mode="f16x3")
        sess = net.stream_session(batch=batch)
        pos = 190 - hop
        o = [sess.feed(x[:, :pos].cuda()), sess.feed(x[:, pos:pos + n].cuda())]
        pos += n
        if mode == "graph":
            sess.capture_graph(n)
        for _ in range(4):
            o.append(sess.feed(x[:, pos:pos + n].cuda())); pos += n
        o.append(sess.feed(x[:, pos:pos + 3 * hop].cuda())); pos += 3 * hop      # other size: eager fallback
        if mode == "graph":
            assert sess._graph is None
        o.append(sess.feed(x[:, pos:pos + n].cuda())); pos += n
        if mode == "graph":
            sess.capture_graph(n)
        for _ in range(3):
            o.append(sess.feed(x[:, pos:pos + n].cuda())); pos += n
        outs[mode] = torch.cat(o, 1).cpu()
        assert pos == x.shape[1]
    assert outs["graph"].shape == outs["eager"].shape and outs["graph"].shape[1] > 0
    assert torch.equal(outs["graph"], outs["eager"])


def test_module_feed_flush_api():
    """Same call pattern as the reference self-test (CleanUMamba.py:574-582): feed + flush ~= parallel forward."""
    fx = load_golden("mini_mamba_442k")
    net = build(fx, normalize_input=False)
    x = fx["noisy"][:1, 0, :8000].cuda()
    with torch.no_grad():
        par = net(x.clone()).squeeze(1)[:, : x.shape[1]]     # normalize_input=False returns the padded length
        seq = net.feed(x)
        seq = torch.cat([seq, net.flush()], 1)
    assert seq.shape == par.shape
    n_exact = seq.shape[1] - 2 * net.total_stride - net.frame_length
    assert (seq[:, :n_exact] - par[:, :n_exact]).abs().max().item() < TOL
    assert torch.allclose(seq, par, atol=0.1)                   # the reference's own tolerance
    with pytest.raises(ValueError):
        net.feed(torch.zeros(2, 3, device="cuda"))


def test_module_feed_hop_by_hop_auto_graph():
    """The reference's real-time loop (examples/streaming_demo.py:118-172): net.feed() one hop at a time.  After a few identical
    hops the step is replayed from a CUDA graph; the stream must equal offline forward and a plain eager session bit for bit."""
    fx = load_golden("e6_pruned_200k")
    net = build(fx, normalize_input=True, math_mode="f16x3")
    hop = net.total_stride
    x = (fx["noisy"][:1, 0, : net.frame_length - hop + 40 * hop] * 1.0).cuda()
    first = net.frame_length - hop
    outs = [net.feed(x[:, :first])]
    for i in range(40):
        outs.append(net.feed(x[:, first + i * hop: first + (i + 1) * hop]))
        if i == 12:
            assert net._stream._graph is not None           # captured by now
    got = torch.cat(outs, 1)
    ref_sess = net.stream_session(batch=1)                  # eager, no auto graph
    want = torch.cat([ref_sess.feed(x[:, :first])] + [ref_sess.feed(x[:, first + i * hop: first + (i + 1) * hop]) for i in range(40)], 1)
    assert got.shape == want.shape == (1, 40 * hop)
    assert torch.equal(got, want)
    tail = net.flush()                                       # releases the graph, clears the conv caches
    assert net._stream._graph is None and tail.shape[1] == first


def test_stream_edge_cases_empty_and_tiny_feeds():
    """feed() with fewer samples than a frame returns (B, 0); one sample at a time still converges to the same output."""
    fx = load_golden("tiny_equalwidth_seed0")
    net = build(fx, normalize_input=False)
    x = fx["noisy"][:1, 0, :400]
    so = orc.StreamOracle(fx["state_dict"], normalize_input=False)
    want = so.feed(x)
    sess = net.stream_session(batch=1)
    assert sess.feed(x[:, :10].cuda()).shape == (1, 0)
    outs = [sess.feed(x[:, 10:11].cuda())]
    outs += [sess.feed(x[:, i:i + 7].cuda()) for i in range(11, x.shape[1], 7)]
    got = torch.cat(outs, 1).cpu()
    assert got.shape == want.shape and (got - want).abs().max().item() < TOL
    with pytest.raises(ValueError):
        sess.feed(torch.zeros(2, 5, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sess.feed(torch.zeros(1, 5))


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("math", ["f16x3", "fp32"])
def test_e6_high_full_size_streaming_matches_stream_oracle(math, normalize):
    """BASELINE configs[2]'s model -- E6 high FULL size (27.2 M, d_inner 2048, d_state 64; seeded random init == reference
    constructor) -- 2 streams, multi-hop chunks (batch-mode GEMMs with >= 128 rows per stream, chunked scan) and single-hop
    chunks (flattened GEMMs, state-update scan at d = 2048), against one StreamOracle per stream in both std modes."""
    from cleanumamba_b200.network import Net
    import os
    sums = json.load(open(os.path.join(__import__("conftest").GOLDEN, "full_init_seed0_sums.json")))["DNS-CleanUMamba-3N-E6"]
    torch.manual_seed(0)
    net = Net("CleanUMamba", dict(sums["config"], math_mode=math, normalize_input=normalize))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().eval()
    B, hop = 2, net.total_stride
    assert hop == 64 and net.frame_length == 190
    sizes = [190 + hop * 39, hop] + [hop] * 5 + [hop * 4, hop * 40, hop, hop * 2, 37, hop - 37]
    g = torch.Generator().manual_seed(33)
    x = torch.randn(B, sum(sizes), generator=g) * 0.1 * (1 + torch.arange(B)[:, None])
    sess = net.stream_session(batch=B)
    outs, pos = [], 0
    for n in sizes:
        outs.append(sess.feed(x[:, pos:pos + n].cuda()))
        pos += n
    outs.append(sess.flush())
    got = torch.cat(outs, 1).cpu()
    worst = 0.0
    for b in range(B):
        so = orc.StreamOracle(sd, normalize_input=normalize)
        want = torch.cat([so.feed(x[b:b + 1]), so.flush()], 1)
        assert got[b:b + 1].shape == want.shape
        worst = max(worst, (got[b:b + 1] - want).abs().max().item())
    print(f"\n[E6-high full streaming {math} normalize={normalize}] max-abs vs StreamOracle {worst:.3e}")
    assert worst <= 1e-4


def test_stream_session_follows_weight_updates_and_frame_reset():
    """A session that already streamed (and captured a graph) must see new parameter values (load_state_dict / optimiser step
    repack the weights: stale graphs are dropped), and ``reset_time_per_frame`` restarts the running input std like the
    reference (:326-328, :399-401).  The conv caches are flushed after the weight change (product and reference carry the
    decoder overlap in different but equivalent forms -- g[p-1] vs the bias-free output tail -- which only coincide for
    unchanged weights); the Mamba state and the running std survive the flush in both."""
    fx = load_golden("e6_pruned_200k")
    net = build(fx, normalize_input=True)
    hop = 64
    g = torch.Generator().manual_seed(8)
    x = torch.randn(1, 190 + hop * 6 + 190 + hop * 5, generator=g) * 0.1
    n1 = 190 + hop * 6
    net.feed(x[:, :190].cuda())
    for k in range(6):          # identical whole-hop chunks: auto graph capture kicks in
        net.feed(x[:, 190 + hop * k: 190 + hop * (k + 1)].cuda())
    assert net._stream._graph is not None
    sd2 = {k: (v * 1.01 if v.dtype.is_floating_point else v) for k, v in fx["state_dict"].items()}
    net.load_state_dict(sd2)
    tail = net.flush().cpu()
    assert net._stream._graph is None and net._stream._pack_gen == net.engine().pack_generation
    net.reset_time_per_frame()
    assert net._stream.frames == 0 and net.frames == 0
    got = net.feed(x[:, n1:].cuda()).cpu()
    so = orc.StreamOracle(fx["state_dict"], normalize_input=True)
    so.feed(x[:, :n1])
    so.sd = {k: v.float() for k, v in sd2.items()}
    tail_want = so.flush()
    so.frames = 0
    want = so.feed(x[:, n1:])
    assert tail.shape == tail_want.shape and (tail - tail_want).abs().max().item() < TOL
    assert got.shape == want.shape and got.shape[1] >= hop * 5
    assert (got - want).abs().max().item() < TOL
