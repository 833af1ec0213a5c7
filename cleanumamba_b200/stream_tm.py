"""Time-major streaming session: ``feed`` for MANY concurrent streams (BASELINE.json configs[2]).

Reference behaviour: ``feed / _denoise_frame / flush`` of /root/reference/src/network/CleanUMamba.py:358-490 -- one stream, one
hop per Python iteration.  ``streaming.StreamSession`` batches streams with every carried buffer laid out (stream, column,
channel); for a few rows per stream and call that needs a gather copy in front of every strided conv, a scatter copy behind
every GEMM that appends to a FIFO, one junk row per stream and level, and ~30 torch ``cat / copy_ / clone`` kernels per call
(~1 ms of a 5.6 ms one-hop call at 4096 streams).

Here every carried buffer is **(column, stream, channel)** -- one contiguous *plane* of all streams per time column:

  * every GEMM of a call has the streams in its M dimension: full 128-row tiles at any number of hops per call, no junk rows;
  * a FIFO is appended to by a plain dense GEMM that writes whole planes at its end (the GLU 1x1 conv of a level writes
    straight into the level's skip FIFO; the decoder's GLU writes behind the carried column of its own FIFO);
  * the strided conv reads its (2 C)-wide input row as two neighbouring planes, the transposed conv its two taps as
    neighbouring planes and writes even / odd output columns as separate planes -- the plane-major operand mode of the
    tap-GEMM (``cum_gemm_desc.a_planes`` / ``n_half``), i.e. no gather, no scatter;
  * the U-Net skip is read in place from the encoder FIFO as the GEMM's addend;
  * ONE ``cum_stream_shift_fwd`` launch at the end of the call moves every FIFO's unconsumed tail to its front.

Channel counts that are not multiples of 32 (pruned checkpoints) run on planes padded to a pitch of 32 with padded copies of the
two convolutions' weights (``engine._pack``: ``enc{i}.wq`` / ``dec{j}.wq``).  All buffers are static (grow-only) so a steady-state call allocates nothing that outlives it, touches only fixed addresses and
is captured as a CUDA graph without any copy-back epilogue.  The arithmetic per stream is the one of ``StreamSession``
(same kernels, same products, same accumulation order): outputs are bit-identical to it.

``state_dtype=torch.float16`` selects the REDUCED-PRECISION variant (reported separately): the carried SSM state, whose
read + write is the HBM floor of a one-hop call, is stored as fp16 (recurrence still fp32).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import EPI_GLU, EPI_RELU, ShiftEntry, ptr
from .streaming import StreamSession, _on_device


def time_major_supported(model) -> bool:
    """Every fp32-storage math mode; the bf16-storage variant is offline-forward only."""
    eng = model.engine()
    eng.ensure_packed()
    return not eng.bf16_io


class TimeMajorStreamSession(StreamSession):
    def __init__(self, model, batch: int = 1, auto_graph: bool = False, state_dtype: torch.dtype = torch.float32):
        if state_dtype not in (torch.float32, torch.float16):
            raise ValueError("state_dtype must be torch.float32 or torch.float16")
        super().__init__(model, batch=batch, auto_graph=auto_graph)       # raises for the bf16-storage variant
        self.state_dtype = state_dtype
        if state_dtype == torch.float16:
            mm0 = self.eng.meta["mamba"]
            if any(mm["N_p"] != 64 or mm["di_p"] % 16 for mm in mm0):
                raise NotImplementedError("cleanumamba_b200: the fp16 SSM state needs d_state = 64 and d_inner % 16 == 0")
            self.states = [(cs, hs.to(torch.float16)) for cs, hs in self.states]

    # ------------------------------------------------------------------------------------------------------ buffers
    def _reset_conv_state(self):
        D, meta, B, dev = self.D, self.eng.meta, self.B, self.dev
        self.enc_base = [0] * D
        self.enc_count = [0] * D
        self.samples_base = 0
        self.frames_since_reset = 0
        # encoder level i: planes [0, enc_count[i] - enc_base[i]) of enc_fifo[i] are the outputs the decoder has not consumed yet
        if not hasattr(self, "enc_fifo"):
            # channel pitch of a plane = q32(C): a plane-major GEMM operand is a whole number of 32-element K-blocks wide.  The pad lanes
            # are zero-initialised and never written (the GEMMs that fill the FIFOs write C_pad columns), the padded weight copies
            # (engine._pack: enc{i}.wq / dec{j}.wq) have zeros there as well
            self.enc_fifo = [torch.zeros(4, B, e["Hoq"], dtype=torch.float32, device=dev) for e in meta["enc"]]
            # decoder level j: plane 0 = the carried GLU column g[-1] (replaces the overlap-add tail, :476-484), planes 1.. = this call's
            self.dec_fifo = [torch.zeros(2, B, d["Hgq"], dtype=torch.float32, device=dev) for d in meta["dec"]]
            self.x_buf = torch.zeros(B, max(4, self.frame_length), dtype=torch.float32, device=dev)
            self.n_pend = 0
        else:
            for g in self.dec_fifo:
                g[0].zero_()
        self._graph = None

    def _grow_planes(self, buf: torch.Tensor, planes: int, keep: int) -> torch.Tensor:
        if buf.shape[0] >= planes:
            return buf
        new = torch.zeros(max(planes, 2 * buf.shape[0]), *buf.shape[1:], dtype=buf.dtype, device=buf.device)
        if keep:
            new[:keep].copy_(buf[:keep])
        self._graph = None          # a captured graph has the old address baked in
        return new

    @property
    def pending(self):
        return self.x_buf[:, : self.n_pend]

    @pending.setter
    def pending(self, value):       # the base constructor assigns an empty tensor
        if value.shape[1] != 0:
            raise RuntimeError("TimeMajorStreamSession.pending is a view of a static buffer")
        self.n_pend = 0

    def pending_view(self):
        return self.x_buf[:, : self.n_pend]

    # ------------------------------------------------------------------------------------------------------ feed
    def _feed(self, chunk: torch.Tensor) -> torch.Tensor:
        n = chunk.shape[1]
        if self._graph is not None:
            if n == self._graph["chunk"]:
                self.x_buf[:, self.n_pend: self.n_pend + n].copy_(chunk)
                self._graph["graph"].replay()
                self._advance(self._graph["F"])
                return self._graph["out"].clone()
            self._graph = None
        if self.auto_graph:
            self._same = self._same + 1 if n == self._last_n else 1
            self._last_n = n
            if self._same > self.AUTO_GRAPH_AFTER and n > 0 and n % self.hop == 0 and self._steady():
                try:
                    self.capture_graph(n)
                    return self._feed(chunk)
                except RuntimeError:
                    self.auto_graph = False
        return self._feed_eager(chunk)

    def _steady(self) -> bool:
        return self.frames_since_reset > 0 and self.n_pend == self.frame_length - self.hop

    def _feed_eager(self, chunk: torch.Tensor) -> torch.Tensor:
        n = chunk.shape[1]
        need = self.n_pend + n
        if need > self.x_buf.shape[1]:
            new = torch.zeros(self.B, (max(need, 2 * self.x_buf.shape[1]) + 3) & ~3, dtype=torch.float32, device=self.dev)
            new[:, : self.n_pend].copy_(self.x_buf[:, : self.n_pend])
            self.x_buf, self._graph = new, None
        if n:
            self.x_buf[:, self.n_pend: need].copy_(chunk)
        if need < self.frame_length:
            self.n_pend = need
            return torch.zeros(self.B, 0, dtype=torch.float32, device=self.dev)
        F = (need - self.frame_length) // self.hop + 1
        out = self._process(F, need)
        self._advance(F, need)
        return out

    def _advance(self, F: int, n_total: int = None):
        """Host-side counters of a processed call of F frames (the device side was done by _process or by the graph replay)."""
        if n_total is None:
            n_total = self.n_pend + self._graph["chunk"]
        D = self.D
        cols = self._new_columns(F)
        for i in range(D):
            self.enc_count[i] += cols[i]
        d = F
        self.enc_base[D - 1] += F
        for j in range(D - 1):
            self.enc_base[D - 2 - j] += 2 * d
            d *= 2
        self.n_pend = n_total - F * self.hop
        self.samples_base += F * self.hop
        self.frames += F
        self.frames_since_reset += F

    def _new_columns(self, F: int):
        """Columns every encoder level produces for a call that completes F frames."""
        n_use = self.frame_length + (F - 1) * self.hop
        avail, cols = self.samples_base + n_use, []
        for i in range(self.D):
            c_new = (avail - 4) // 2 + 1
            cols.append(c_new - self.enc_count[i])
            avail = c_new
        return cols

    # ------------------------------------------------------------------------------------------------------ graph
    _COUNTERS = StreamSession._COUNTERS + ("n_pend",)

    @torch.no_grad()
    @_on_device
    def capture_graph(self, chunk_samples: int) -> None:
        """Capture feed() for chunks of exactly ``chunk_samples`` samples (a multiple of the hop) in steady state.  Every buffer of
        the session is static, so the graph is the plain launch sequence of one call (no copy-back epilogue); other chunk sizes
        fall back to eager."""
        if chunk_samples <= 0 or chunk_samples % self.hop:
            raise ValueError(f"chunk_samples must be a positive multiple of the hop ({self.hop})")
        if not self._steady():
            raise RuntimeError("capture_graph: the session is not in steady state (feed at least one full frame first, and "
                               "feed whole hops so that pending holds frame_length - hop samples)")
        self._sync_weights()
        self._graph = None
        if self.model.normalize_input and self.frames_dev is None:
            self.frames_dev = torch.tensor([self.frames], dtype=torch.int32, device=self.dev)
        F = chunk_samples // self.hop
        n_total = self.n_pend + chunk_samples
        # one eager step on a saved copy of ALL state: grows every buffer to its steady-state size and runs every lazy
        # initialisation of this exact launch sequence (tensor maps, kernel attributes) outside the capture
        snap = self._snapshot()
        fill = [self.enc_count[i] - self.enc_base[i] for i in range(self.D)]
        saved = dict(x=self.x_buf[:, : self.n_pend].clone(), enc=[f[:k].clone() for f, k in zip(self.enc_fifo, fill)],
                     dec=[g[0].clone() for g in self.dec_fifo], std=self.running_std.clone(),
                     st=[(a.clone(), b.clone()) for a, b in self.states],
                     fd=None if self.frames_dev is None else self.frames_dev.clone())
        self._feed_eager(torch.zeros(self.B, chunk_samples, dtype=torch.float32, device=self.dev))
        after = [self.enc_count[i] - self.enc_base[i] for i in range(self.D)]
        self._restore(snap)
        self.x_buf[:, : self.n_pend].copy_(saved["x"])
        for f, s in zip(self.enc_fifo, saved["enc"]):
            f[: s.shape[0]].copy_(s)
        for g, s in zip(self.dec_fifo, saved["dec"]):
            g[0].copy_(s)
        self.running_std.copy_(saved["std"])
        for (a, b), (sa, sb) in zip(self.states, saved["st"]):
            a.copy_(sa)
            b.copy_(sb)
        if self.frames_dev is not None:
            self.frames_dev.copy_(saved["fd"])
        self._graph = None
        if after != fill:
            raise RuntimeError("capture_graph: FIFO fill changes from step to step (not in steady state for this chunk size)")
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._process(F, n_total)
        self._graph = dict(graph=graph, chunk=chunk_samples, F=F, out=out)

    def release_graph(self) -> None:
        self._graph = None

    @torch.no_grad()
    @_on_device
    def flush(self) -> torch.Tensor:
        """:358-368 -- clear the conv caches, feed frame_length zeros, return the first len(pending) samples."""
        n = self.n_pend
        self._reset_conv_state()            # keeps the pending samples, the Mamba state and the running std
        out = self.feed(torch.zeros(self.B, self.frame_length, dtype=torch.float32, device=self.dev))
        return out[:, :n]

    # ------------------------------------------------------------------------------------------------------ one call
    def _process(self, F: int, n_total: int) -> torch.Tensor:
        """Device side of one call: F complete frames out of the first ``n_total`` samples of x_buf.  Reads the host counters,
        does not change them (``_advance`` does)."""
        eng, m, B, D, dev = self.eng, self.model, self.B, self.D, self.dev
        pk, meta, lib = eng.pk, eng.meta, eng.lib
        act = EPI_GLU[m.glu_activation]
        st = _lib.stream_ptr
        X = self.x_buf
        xs = X.shape[1]
        n_use = self.frame_length + (F - 1) * self.hop
        first = self.frames_since_reset == 0
        shifts = []                                   # (tensor, row_stride, src_off, count, rows): end-of-call FIFO maintenance

        scale = None
        if m.normalize_input:
            scale = torch.empty(B, F, dtype=torch.float32, device=dev)
            if self.frames_dev is not None:
                eng._call("stream_std", lib.cum_stream_std_counter_fwd, X.data_ptr(), xs, B, F, self.frame_length, self.hop,
                          self.frames_dev.data_ptr(), self.running_std.data_ptr(), scale.data_ptr(), st(), launches=2)
            else:
                eng._call("stream_std", lib.cum_stream_std_fwd, X.data_ptr(), xs, B, F, self.frame_length, self.hop,
                          self.frames, self.running_std.data_ptr(), scale.data_ptr(), st())

        # ---------------- encoder
        cols = self._new_columns(F)
        fill = [self.enc_count[i] - self.enc_base[i] for i in range(D)]       # valid planes per FIFO before this call
        for i, e in enumerate(meta["enc"]):
            rows_new, c_old = cols[i], self.enc_count[i]
            assert rows_new > 0
            y = torch.empty(rows_new, B, e["Hc_p"], dtype=torch.float32, device=dev)
            if i == 0:
                off = 2 * c_old - self.samples_base
                per_frame = self.hop // 2
                first_rows = (self.frame_length - 4) // 2 + 1
                row_off = -(first_rows - per_frame) if first else 0
                eng._call("conv_in", lib.cum_conv_in_strided_fwd, X.data_ptr() + 4 * off, xs, B, n_use - off,
                          pk["enc0.w"].data_ptr(), pk["enc0.b"].data_ptr(), y.data_ptr(), e["Hc_p"], B * e["Hc_p"], rows_new, e["Hc_p"], 4, 2,
                          ptr(scale), per_frame, row_off, st())
            else:
                src, cq = self.enc_fifo[i - 1], meta["enc"][i - 1]["Hoq"]
                wkey = f"enc{i}.wq" if f"enc{i}.wq" in pk else f"enc{i}.w"
                lo = 2 * c_old - self.enc_base[i - 1]                    # plane of the first input column this call needs
                # output column t of every stream = W01 . [plane lo+2t | plane lo+2t+1] + W23 . [plane lo+2t+2 | lo+2t+3]
                eng.gemm(src, 0, B * cq, cq, B, 2 * cq, wkey, pk[f"enc{i}.b"], y, 0, B * e["Hc_p"], e["Hc_p"], B, e["Hc_p"],
                         rows_new, EPI_RELU, taps=2, shifts=(0, 1), planes=(fill[i - 1] + cols[i - 1], cq, lo, 2, 0), small=self._small(rows_new, f"enc{i}.w"))
            ho, hoq = e["Ho_p"], e["Hoq"]
            self.enc_fifo[i] = self._grow_planes(self.enc_fifo[i], fill[i] + rows_new, fill[i])
            # the GLU 1x1 conv appends its planes to the level's FIFO directly
            eng.dense(y, rows_new * B, e["Hc_p"], f"enc{i}.wg", pk[f"enc{i}.bg"], 2 * ho, epi=act, out=self.enc_fifo[i],
                      out_off=fill[i] * B * hoq, out_rs=hoq, small=self._small(rows_new, f"enc{i}.wg"))

        # ---------------- bottleneck: F tokens per stream, time-major rows (t * B + b)
        last = self.enc_fifo[D - 1]
        cbp, cbq = meta["enc"][-1]["Ho_p"], meta["enc"][-1]["Hoq"]
        assert fill[D - 1] + cols[D - 1] == F, (fill[D - 1], cols[D - 1], F)
        h = eng.dense(last, B * F, cbp, "t1.w", pk["t1.b"], meta["dm_p"], a_rs=cbq, small=self._small(F, "t1.w"))
        hn = eng.mamba_layers(h, B, F, states=self.states, tm=True, small=self._small(F, "m0.in"))
        xcur = eng.dense(hn, B * F, meta["dm_p"], "t2.w", pk["t2.b"], cbp, addend=last, add_rs=cbq, small=self._small(F, "t2.w"))
        x_rs = cbp                                      # row pitch of xcur

        # ---------------- decoder
        d_cols = F
        out = None
        for j, dd in enumerate(meta["dec"]):
            hg, hgq = dd["Hg_p"], dd["Hgq"]
            self.dec_fifo[j] = G = self._grow_planes(self.dec_fifo[j], d_cols + 1, 1)
            eng.dense(xcur, B * d_cols, dd["Cin_p"], f"dec{j}.wg", pk[f"dec{j}.bg"], 2 * hg, epi=act, out=G, out_off=B * hgq, out_rs=hgq,
                      a_rs=x_rs, small=self._small(d_cols, f"dec{j}.wg"))
            shifts.append((G, 0, d_cols * B * hgq, B * hgq, 1))          # the last column becomes the carried one
            if j < D - 1:
                coq = dd["Coq"]
                lvl = D - 2 - j
                skip, skq = self.enc_fifo[lvl], meta["enc"][lvl]["Hoq"]
                wkey, bkey = (f"dec{j}.wq", f"dec{j}.bq") if f"dec{j}.wq" in pk else (f"dec{j}.w", f"dec{j}.b")
                nxt = torch.empty(2 * d_cols, B, coq, dtype=torch.float32, device=dev)
                # output column 2p + par of every stream = Wa_par . G[p + 1] + Wb_par . G[p] + skip column 2p + par (read in place)
                eng.gemm(G, 0, B * hgq, hgq, B, hgq, wkey, pk[bkey], nxt, 0, B * coq, coq, B, coq, 2 * d_cols, EPI_RELU,
                         taps=2, shifts=(1, 0), addend=skip, add_bs=B * skq, add_rs=skq, planes=(d_cols + 1, hgq, 0, 1, 1),
                         small=self._small(2 * d_cols, f"dec{j}.w", half=True))
                xcur, x_rs = nxt, coq
                d_cols = 2 * d_cols
            else:
                length = 2 * d_cols
                out = torch.empty(B, length, dtype=torch.float32, device=dev)
                eng._call("convt_out", lib.cum_convt_out_strided_fwd, G.data_ptr(), hgq, B * hgq, B, d_cols + 1, hg,
                          pk[f"dec{j}.w"].data_ptr(), meta["out_bias"], ptr(scale), self.hop, out.data_ptr(), length, 2, length, 4, 2, st())

        # ---------------- FIFO maintenance: one launch
        d = F
        consumed = [0] * D
        consumed[D - 1] = F
        for j in range(D - 1):
            consumed[D - 2 - j] = 2 * d
            d *= 2
        for i, e in enumerate(meta["enc"]):
            left = fill[i] + cols[i] - consumed[i]
            assert left >= 0
            if left:
                pe = B * e["Hoq"]
                shifts.append((self.enc_fifo[i], 0, consumed[i] * pe, left * pe, 1))
        left = n_total - F * self.hop
        if left:
            shifts.append((X, xs, F * self.hop, left, B))
        tab = (ShiftEntry * len(shifts))()
        for k, (t, rs, so, cnt, rows) in enumerate(shifts):
            tab[k].base, tab[k].row_stride, tab[k].src_off, tab[k].count, tab[k].rows = t.data_ptr(), rs, so, cnt, rows
        eng._call("stream_shift", lib.cum_stream_shift_fwd, tab, len(shifts), st())
        return out
