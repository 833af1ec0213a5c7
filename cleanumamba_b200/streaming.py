"""Carried-state streaming inference for a batch of independent streams.

Reference behaviour being reproduced: ``feed / _denoise_frame / flush`` of /root/reference/src/network/CleanUMamba.py:358-490
(one stream, one 2^D-sample hop per Python iteration, ~60 tiny kernels per hop).  Here ``batch`` streams advance in
lock-step and every ``feed`` processes ALL complete frames of the call in one pass through the same kernels as the
offline forward, with the state that makes chunked == offline carried between calls:

  * raw ``pending`` samples (the reference's ``self.pending``)                                  (:392,410)
  * per encoder level, the output columns the decoder has not consumed yet -- they double as the K-S = 2 input columns
    the next strided conv still needs (the reference's ``enc{i}`` caches, :432-442)
  * per decoder level, ONE column of the GLU output: ``convT(g)[2p+par] = Wa g[p] + Wb g[p-1]``, so carrying g[p-1]
    replaces the reference's overlap-add tail ``x[..., -stride:] - bias`` (:476-484) exactly
  * per Mamba layer, conv_state (B, 3, d_inner) and ssm_state (B, d_inner, d_state) updated in place by the
    depthwise-conv / scan kernels (``Mamba.step``), any number of tokens per call
  * the running mean of the per-frame input std and the frame counter (:399-401)

Identity (normalize_input=False): the samples emitted equal offline ``forward`` on the fed signal.  With
``normalize_input=True`` the reference's per-frame running-std rule is applied hop by hop (conv_in divides each new
encoder column by the std of the frame that produced it; the output hop is multiplied by the same value).
The reference's skip-index bug (:474) is NOT reproduced (it crashes on every shipped checkpoint).
"""
from __future__ import annotations

import functools
import os

import torch

from . import _lib
from ._lib import EPI_GLU, EPI_RELU, ptr


def _on_device(fn):
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self.eng.device_guard():
            return fn(self, *args, **kwargs)
    return wrapper


class StreamSession:
    BATCH_MODE_ROWS = 128     # rows per stream from which a level runs one GEMM batch item per stream (in-place operands)
    AUTO_GRAPH_AFTER = 3      # auto_graph sessions: identical chunks seen before the step is captured as a CUDA graph
    SMALL_STREAMS = int(os.environ.get("CUM_STREAM_SMALL_STREAMS", 4))
    SMALL_MACS = int(float(os.environ.get("CUM_STREAM_SMALL_MMAC", 16)) * 1e6)
    # Sessions of up to SMALL_STREAMS streams run the GEMMs of a call on the CUDA-core small-M path (cum_gemm_desc.small_m_path: ~10 us
    # less fixed cost per launch than the tensor-core pipeline, ~0.4 TMAC/s) as long as a GEMM stays under SMALL_MACS multiply-adds
    # (multi-hop calls of a single stream go back to the tensor cores).  Decided from (columns, streams, weight shape) -- not from the
    # GEMM's own m / batch, which differ between the two buffer layouts -- so that the stream-major and the time-major session run
    # the same kernel for the same level.  Measured per 1-hop call from the graph, all GEMMs small-M vs none: 1 stream 0.32 vs
    # 0.73 ms, 2 streams 0.35 vs 0.73, 4 streams 0.49 vs 0.73, 8 streams 0.83 vs 0.74

    def _small(self, cols: int, w: str, half: bool = False) -> int:
        if self.B > self.SMALL_STREAMS:
            return -1
        macs = cols * self.B * self.eng.pk[w].numel() // (2 if half else 1)
        return 1 if macs <= self.SMALL_MACS else -1

    def __init__(self, model, batch: int = 1, auto_graph: bool = False):
        self.auto_graph, self._same, self._last_n = auto_graph, 0, -1
        self.model = model
        self.eng = model.engine()
        self.eng.ensure_packed()
        self._pack_gen = self.eng.pack_generation
        if self.eng.bf16_io:
            raise NotImplementedError("cleanumamba_b200: math_mode='bf16' (reduced-precision variant) is offline-forward only")
        self.B = batch
        self.dev = self.eng.device
        m, meta = model, self.eng.meta
        self.D = meta["D"]
        self.hop = m.total_stride
        self.frame_length = m.frame_length
        self.frames = 0                      # frames processed since construction (running-std denominator)
        self.running_std = torch.zeros(batch, dtype=torch.float32, device=self.dev)
        self.pending = torch.zeros(batch, 0, dtype=torch.float32, device=self.dev)
        self.states = [(torch.zeros(batch, mm["W"] - 1, mm["di_p"], dtype=torch.float32, device=self.dev),
                        torch.zeros(batch, mm["di_p"], mm["N_p"], dtype=torch.float32, device=self.dev))
                       for mm in meta["mamba"]]
        self.frames_dev = None               # device copy of `frames` (graph mode: the running-std kernel reads it there)
        self._graph = None                   # (CUDAGraph, chunk_samples, static input, static output, static state, counter deltas)
        self._reset_conv_state()

    # ------------------------------------------------------------------------------------------------------
    def _reset_conv_state(self):
        """Fresh encoder/decoder caches (what ``flush`` clears, :364); Mamba state and running std are kept."""
        D, meta = self.D, self.eng.meta
        self.enc_buf = [None] * D            # (B, cap, C_p): encoder level i outputs from absolute column enc_base[i]
        self.enc_base = [0] * D              # columns already consumed by the decoder
        self.enc_count = [0] * D             # columns produced so far
        self.samples_base = 0                # absolute index (since reset) of pending[:, 0]
        self.frames_since_reset = 0
        self.dec_carry = [torch.zeros(self.B, d["Hg_p"], dtype=torch.float32, device=self.dev) for d in meta["dec"]]

    def pending_view(self):
        return self.pending

    def reset_frames(self):
        """Restart the running mean of the per-frame input std (the reference's ``reset_time_per_frame`` zeroes ``frames``, its
        denominator, :326-328 / :399-401)."""
        self.frames = 0
        if self.frames_dev is not None:
            self.frames_dev.zero_()

    def _sync_weights(self):
        """Parameters may have changed since the last call (load_state_dict, an optimiser step, .half()/.float()): repack, and
        drop a captured graph -- it has the previous packed-weight pointers baked in."""
        self.eng.ensure_packed()
        if self.eng.pack_generation != self._pack_gen:
            self._pack_gen = self.eng.pack_generation
            self.release_graph()

    # ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def feed(self, chunk: torch.Tensor) -> torch.Tensor:
        """chunk: (batch, n) -> (batch, hop * F), F = number of complete frames now available."""
        if chunk.dim() != 2 or chunk.shape[0] != self.B:
            raise ValueError(f"expected a (batch={self.B}, n) tensor, got {tuple(chunk.shape)}")
        if chunk.device != self.dev:
            raise RuntimeError(f"chunk is on {chunk.device}, model on {self.dev} (no CPU fallback)")
        self._sync_weights()
        with self.eng.device_guard():
            return self._feed(chunk)

    def _feed(self, chunk: torch.Tensor) -> torch.Tensor:
        if self._graph is not None:
            if chunk.shape[1] == self._graph["chunk"]:
                return self._replay(chunk)
            self.release_graph()             # a different chunk size: back to the eager path (state stays valid)
        if self.auto_graph:
            # the module-level feed() (one stream, the reference's real-time loop): after AUTO_GRAPH_AFTER identical whole-hop
            # chunks in steady state the step is captured and replayed from then on (bit-identical, ~2.5x lower latency)
            n = chunk.shape[1]
            self._same = self._same + 1 if n == self._last_n else 1
            self._last_n = n
            if (self._same > self.AUTO_GRAPH_AFTER and n > 0 and n % self.hop == 0 and self.frames_since_reset > 0
                    and self.pending.shape[1] == self.frame_length - self.hop):
                try:
                    self.capture_graph(n)
                    return self._replay(chunk)
                except RuntimeError:
                    self.auto_graph = False  # not capturable in this state: stay eager
        return self._feed_eager(chunk)

    def _feed_eager(self, chunk: torch.Tensor) -> torch.Tensor:
        self.pending = torch.cat([self.pending, chunk.to(torch.float32)], dim=1)
        n = self.pending.shape[1]
        if n < self.frame_length:
            return torch.zeros(self.B, 0, dtype=torch.float32, device=self.dev)
        F = (n - self.frame_length) // self.hop + 1
        out = self._process(F)
        self.pending = self.pending[:, F * self.hop:]
        self.samples_base += F * self.hop
        return out

    # ------------------------------------------------------------------------------------------------------
    # Steady-state CUDA graph.  A streaming step is ~260 small launches plus Python glue: launch-bound for one or a few
    # hundred streams (2.1 ms per hop eagerly, regardless of the stream count).  Once the session is in steady state for a
    # fixed chunk size every shape, FIFO offset and operand address of a step is constant, so the whole feed() is captured
    # once and replayed: the carried state lives in static tensors (the graph ends by copying the new FIFO tails / carries
    # back into them), the frame count of the running std lives on the device (cum_stream_std_counter_fwd).
    # ------------------------------------------------------------------------------------------------------
    _COUNTERS = ("enc_base", "enc_count", "samples_base", "frames", "frames_since_reset")

    def _snapshot(self):
        return {k: (list(getattr(self, k)) if isinstance(getattr(self, k), list) else getattr(self, k)) for k in self._COUNTERS}

    def _restore(self, snap):
        for k, v in snap.items():
            setattr(self, k, list(v) if isinstance(v, list) else v)

    @torch.no_grad()
    @_on_device
    def capture_graph(self, chunk_samples: int) -> None:
        """Capture feed() for chunks of exactly ``chunk_samples`` samples (a multiple of the hop).  Call after at least one
        eager feed() has produced output (steady state); feeds of any other size transparently fall back to the eager path."""
        if chunk_samples <= 0 or chunk_samples % self.hop:
            raise ValueError(f"chunk_samples must be a positive multiple of the hop ({self.hop})")
        if self.frames_since_reset == 0 or self.pending.shape[1] != self.frame_length - self.hop:
            raise RuntimeError("capture_graph: the session is not in steady state (feed at least one full frame first, and "
                               "feed whole hops so that pending holds frame_length - hop samples)")
        self._sync_weights()
        self.release_graph()
        if self.model.normalize_input and self.frames_dev is None:
            self.frames_dev = torch.tensor([self.frames], dtype=torch.int32, device=self.dev)
        # static state tensors
        self.pending = self.pending.contiguous().clone()
        self.enc_buf = [b.contiguous().clone() for b in self.enc_buf]
        self.dec_carry = [c.contiguous().clone() for c in self.dec_carry]
        static = dict(pending=self.pending, enc=list(self.enc_buf), carry=list(self.dec_carry))
        gx = torch.zeros(self.B, chunk_samples, dtype=torch.float32, device=self.dev)
        # one eager step on a copy of ALL state, so that every lazy initialisation of this exact launch sequence has happened
        snap = self._snapshot()
        saved = ([t.clone() for t in (static["pending"], *static["enc"], *static["carry"], self.running_std)],
                 [(a.clone(), b.clone()) for a, b in self.states], None if self.frames_dev is None else self.frames_dev.clone())
        self._feed_eager(gx)
        shapes_after = [tuple(b.shape) for b in self.enc_buf]
        after = self._snapshot()
        for dst, src in zip((static["pending"], *static["enc"], *static["carry"], self.running_std), saved[0]):
            dst.copy_(src)
        for (a, b), (sa, sb) in zip(self.states, saved[1]):
            a.copy_(sa); b.copy_(sb)
        if self.frames_dev is not None:
            self.frames_dev.copy_(saved[2])
        if shapes_after != [tuple(b.shape) for b in static["enc"]]:
            self.pending, self.enc_buf, self.dec_carry = static["pending"], list(static["enc"]), list(static["carry"])
            self._restore(snap)
            raise RuntimeError("capture_graph: FIFO shapes change from step to step (not in steady state for this chunk size)")
        self.pending, self.enc_buf, self.dec_carry = static["pending"], list(static["enc"]), list(static["carry"])
        self._restore(snap)
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._feed_eager(gx)
            static["pending"].copy_(self.pending)
            for dst, src in zip(static["enc"], self.enc_buf):
                dst.copy_(src)
            for dst, src in zip(static["carry"], self.dec_carry):
                dst.copy_(src)
        # capture records, it does not execute: the state is still the pre-capture one
        self.pending, self.enc_buf, self.dec_carry = static["pending"], list(static["enc"]), list(static["carry"])
        self._restore(snap)
        delta = {k: ([x - y for x, y in zip(after[k], snap[k])] if isinstance(snap[k], list) else after[k] - snap[k]) for k in snap}
        self._graph = dict(graph=graph, chunk=chunk_samples, x=gx, out=out, static=static, delta=delta)

    def release_graph(self) -> None:
        self._graph = None

    def _replay(self, chunk: torch.Tensor) -> torch.Tensor:
        g = self._graph
        g["x"].copy_(chunk)
        g["graph"].replay()
        for k, d in g["delta"].items():
            cur = getattr(self, k)
            setattr(self, k, [c + x for c, x in zip(cur, d)] if isinstance(cur, list) else cur + d)
        return g["out"].clone()

    @torch.no_grad()
    @_on_device
    def flush(self) -> torch.Tensor:
        """:358-368 -- clear the conv caches, feed frame_length zeros, return the first len(pending) samples."""
        self.release_graph()
        self._reset_conv_state()
        n = self.pending.shape[1]
        out = self.feed(torch.zeros(self.B, self.frame_length, dtype=torch.float32, device=self.dev))
        return out[:, :n]

    # ------------------------------------------------------------------------------------------------------
    def _process(self, F: int) -> torch.Tensor:
        eng, m, B, D, dev = self.eng, self.model, self.B, self.D, self.dev
        pk, meta, lib = eng.pk, eng.meta, eng.lib
        act = EPI_GLU[m.glu_activation]
        st = _lib.stream_ptr
        X = self.pending.contiguous()
        self.pending = X
        n_use = self.frame_length + (F - 1) * self.hop          # samples of X that belong to complete frames
        first = self.frames_since_reset == 0

        scale = None
        if m.normalize_input:
            scale = torch.empty(B, F, dtype=torch.float32, device=dev)
            if self.frames_dev is not None:
                eng._call("stream_std", lib.cum_stream_std_counter_fwd, X.data_ptr(), X.shape[1], B, F, self.frame_length, self.hop,
                          self.frames_dev.data_ptr(), self.running_std.data_ptr(), scale.data_ptr(), st(), launches=2)
            else:
                eng._call("stream_std", lib.cum_stream_std_fwd, X.data_ptr(), X.shape[1], B, F, self.frame_length, self.hop,
                          self.frames, self.running_std.data_ptr(), scale.data_ptr(), st())

        # ---------------- encoder: every level produces all columns its (frame-aligned) input allows.
        # All streams are flattened into the GEMM M dimension (full 128-row MMA tiles even when a stream contributes only
        # a few rows): the strided conv reads a compact per-stream window of P = rows_new + 1 two-column rows, so row
        # b*P + t of the flat problem is output t of stream b and the one junk row per stream (t = rows_new, whose second
        # tap would reach into the next stream) is dropped when the output FIFO is assembled.
        avail_in = self.samples_base + n_use                    # absolute count of level-0 inputs (samples)
        for i, e in enumerate(meta["enc"]):
            c_old = self.enc_count[i]
            c_new = (avail_in - 4) // 2 + 1
            rows_new = c_new - c_old
            assert rows_new > 0
            if i == 0:
                P = rows_new
                y = torch.empty(B * P, e["Hc_p"], dtype=torch.float32, device=dev)
                off = 2 * c_old - self.samples_base             # first sample of the first new column, inside X
                per_frame = self.hop // 2                       # new level-1 columns per frame
                first_rows = (self.frame_length - 4) // 2 + 1   # ... of the first frame after a reset
                row_off = -(first_rows - per_frame) if first else 0
                eng._call("conv_in", lib.cum_conv_in_fwd, X.data_ptr() + 4 * off, X.shape[1], B, n_use - off,
                          pk["enc0.w"].data_ptr(), pk["enc0.b"].data_ptr(), y.data_ptr(), rows_new, e["Hc_p"], 4, 2,
                          ptr(scale), per_frame, row_off, st())
            elif rows_new >= self.BATCH_MODE_ROWS:
                # enough rows per stream for full 128-row tiles: one GEMM batch item per stream, operands addressed IN PLACE
                # (the input window inside the previous level's FIFO, the output slot inside this level's FIFO): no staging copies
                P = rows_new
                src, cp = self.enc_buf[i - 1], e["Cin_p"]
                lo = 2 * c_old - self.enc_base[i - 1]
                y = torch.empty(B * P, e["Hc_p"], dtype=torch.float32, device=dev)
                eng.gemm(src, lo * cp, src.stride(0), 2 * cp, (src.shape[1] - lo) // 2, 2 * cp, f"enc{i}.w", pk[f"enc{i}.b"],
                         y, 0, P * e["Hc_p"], e["Hc_p"], P, e["Hc_p"], B, EPI_RELU, taps=2, shifts=(0, 1))
            else:
                P = rows_new + 1
                src, cp = self.enc_buf[i - 1], e["Cin_p"]
                lo = 2 * c_old - self.enc_base[i - 1]           # local column of the first input this call needs
                cin = src[:, lo: lo + 2 * P].contiguous()       # (B, 2P, cp) == B*P rows of 2*cp
                y = torch.empty(B * P, e["Hc_p"], dtype=torch.float32, device=dev)
                eng.gemm(cin, 0, 0, 2 * cp, B * P, 2 * cp, f"enc{i}.w", pk[f"enc{i}.b"],
                         y, 0, 0, e["Hc_p"], B * P, e["Hc_p"], 1, EPI_RELU, taps=2, shifts=(0, 1), small=self._small(rows_new, f"enc{i}.w"))
            # output FIFO of this level: [columns the decoder has not consumed yet | new columns]
            keep = c_old - self.enc_base[i]
            buf = torch.empty(B, keep + rows_new, e["Ho_p"], dtype=torch.float32, device=dev)
            if keep:
                old = self.enc_buf[i]
                buf[:, :keep].copy_(old[:, old.shape[1] - keep:])
            if rows_new >= self.BATCH_MODE_ROWS:       # batch mode: the GLU GEMM writes straight into the FIFO slot
                ho = e["Ho_p"]
                eng.gemm(y, 0, P * e["Hc_p"], e["Hc_p"], P, e["Hc_p"], f"enc{i}.wg", pk[f"enc{i}.bg"],
                         buf, keep * ho, (keep + rows_new) * ho, ho, rows_new, 2 * ho, B, act)
            else:
                newc = eng.dense(y, B * P, e["Hc_p"], f"enc{i}.wg", pk[f"enc{i}.bg"], 2 * e["Ho_p"], epi=act, small=self._small(rows_new, f"enc{i}.wg"))
                buf[:, keep:].copy_(newc.view(B, P, e["Ho_p"])[:, :rows_new])
            self.enc_buf[i] = buf
            self.enc_count[i] = c_new
            avail_in = c_new

        # ---------------- bottleneck: F new tokens
        last = self.enc_buf[D - 1]
        cbp = meta["enc"][-1]["Ho_p"]
        assert last.shape[1] == F, (last.shape, F)
        h = eng.dense(last, B * F, cbp, "t1.w", pk["t1.b"], meta["dm_p"], small=self._small(F, "t1.w"))
        hn = eng.mamba_layers(h, B, F, states=self.states, small=self._small(F, "m0.in"))
        xcur = eng.dense(hn, B * F, meta["dm_p"], "t2.w", pk["t2.b"], cbp, addend=last, small=self._small(F, "t2.w"))
        self.enc_base[D - 1] += F
        self.enc_buf[D - 1] = last[:, F:]

        # ---------------- decoder: d columns in, 2d final columns out per level (same flattening, pitch d + 1)
        d_cols = F
        out = None
        for j, dd in enumerate(meta["dec"]):
            hg = dd["Hg_p"]
            batch_mode = d_cols >= self.BATCH_MODE_ROWS
            G = torch.empty(B, d_cols + 1, hg, dtype=torch.float32, device=dev)
            G[:, 0].copy_(self.dec_carry[j])                    # carried g[-1]: replaces the overlap-add tail (:476-484)
            if batch_mode:      # one batch item per stream: the GLU output lands in G[:, 1:] directly
                eng.gemm(xcur, 0, d_cols * dd["Cin_p"], dd["Cin_p"], d_cols, dd["Cin_p"], f"dec{j}.wg", pk[f"dec{j}.bg"],
                         G, hg, (d_cols + 1) * hg, hg, d_cols, 2 * hg, B, act)
            else:
                gnew = eng.dense(xcur, B * d_cols, dd["Cin_p"], f"dec{j}.wg", pk[f"dec{j}.bg"], 2 * hg, epi=act, small=self._small(d_cols, f"dec{j}.wg"))
                G[:, 1:].copy_(gnew.view(B, d_cols, hg))
            self.dec_carry[j] = G[:, d_cols].clone()
            if j < D - 1 and batch_mode:
                co = dd["Co_p"]
                lvl = D - 2 - j
                skip = self.enc_buf[lvl]
                nxt = torch.empty(B * d_cols, 2 * co, dtype=torch.float32, device=dev)
                # row p of stream b = Wa . G[b, p+1] + Wb . G[b, p] + skip columns (2p, 2p+1), read in place from the encoder FIFO
                eng.gemm(G, 0, (d_cols + 1) * hg, hg, d_cols + 1, hg, f"dec{j}.w", pk[f"dec{j}.b"], nxt, 0, d_cols * 2 * co, 2 * co,
                         d_cols, 2 * co, B, EPI_RELU, taps=2, shifts=(1, 0), addend=skip, add_bs=skip.stride(0), add_rs=2 * co)
                self.enc_base[lvl] += 2 * d_cols
                self.enc_buf[lvl] = skip[:, 2 * d_cols:]
                xcur = nxt                                       # (B, d_cols, 2 co) == (B * 2 d_cols, co)
                d_cols = 2 * d_cols
            elif j < D - 1:
                co = dd["Co_p"]
                lvl = D - 2 - j
                skip = self.enc_buf[lvl]
                Pd = d_cols + 1
                skipc = torch.empty(B, Pd, 2 * co, dtype=torch.float32, device=dev)
                skipc[:, :d_cols].copy_(skip[:, : 2 * d_cols].reshape(B, d_cols, 2 * co))
                nxt = torch.empty(B * Pd, 2 * co, dtype=torch.float32, device=dev)
                # flat row b*Pd + p = Wa . G[b, p+1] + Wb . G[b, p]; p = d_cols is the junk row of stream b
                eng.gemm(G, 0, 0, hg, B * Pd, hg, f"dec{j}.w", pk[f"dec{j}.b"], nxt, 0, 0, 2 * co, B * Pd, 2 * co, 1,
                         EPI_RELU, taps=2, shifts=(1, 0), addend=skipc, add_bs=0, add_rs=2 * co, small=self._small(2 * d_cols, f"dec{j}.w", half=True))
                self.enc_base[lvl] += 2 * d_cols
                self.enc_buf[lvl] = skip[:, 2 * d_cols:]
                xcur = nxt.view(B, Pd, 2 * co)[:, :d_cols].reshape(B * 2 * d_cols, co)
                d_cols = 2 * d_cols
            else:
                length = 2 * d_cols
                out = torch.empty(B, length, dtype=torch.float32, device=dev)
                eng._call("convt_out", lib.cum_convt_out_fwd, G.data_ptr(), B, d_cols + 1, hg, pk[f"dec{j}.w"].data_ptr(),
                          meta["out_bias"], ptr(scale), self.hop, out.data_ptr(), length, 2, length, 4, 2, st())
        self.frames += F
        self.frames_since_reset += F
        return out
