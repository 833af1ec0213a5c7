#!/usr/bin/env python
"""bench.py -- headline benchmark of the CleanUMamba hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--seconds S] [--model e8|e6]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): CleanUMamba E8 full (41.37 M params, seeded random init -- the full checkpoints are
not shipped), batch 64 x 10 s of 16 kHz synthetic noisy speech per GPU, offline forward, fp32 storage and
fp32-tolerance arithmetic (default math mode f16x3; --math fp32 = exact FFMA).  A "step" is
one forward pass over one batch.  Metric: audio-seconds denoised per wall-second, aggregate over all GPUs (utterance
sharding, no data-path collective -> weak scaling).

Own arm     : the product (cleanumamba_b200 -> libcleanumamba_sm100.so).  `value` = inputs resident in HBM, CUDA-event
              timed, max over ranks; `e2e` = same call with pinned HOST buffers, H2D + forward + D2H inside the timing.
Reference arm (--impl reference): the reference's own CPU implementation of the path on all host cores, each step a bounded
              sample (1 clip) of the same workload: the UNMODIFIED reference files (src/network/CleanUMamba.py ... staged
              byte-for-byte into the git-ignored baseline/_ref/ by build(), imported through oracle/ref_loader.py on the
              mamba_ssm stand-in oracle/ref_shim: selective_scan_ref) -> kind "reference"; the oracle port
              (oracle/cleanumamba_oracle.py) is timed beside it (`port`) and is the fallback when baseline/_ref is absent.
Extras      : the default line also carries `train` (configs[3]: DP training step, NCCL gradient all-reduce at N > 1, exposed
              all-reduce time) and `stream` (configs[2]: 4096 streams in total, sharded over the ranks) sub-records, so the
              driver's 1/2/4/8-GPU scaling run measures the collective path too (--no-extras skips them).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "audio-sec denoised/sec (16 kHz)"
UNIT = "audio-s/s"
SR = 16000
CONFIGS = {
    "e8": dict(channels_input=1, channels_output=1, channels_H=64, max_H=768, encoder_n_layers=8, kernel_size=4,
               stride=2, tsfm_n_layers=3, tsfm_n_head=8, tsfm_d_model=512, tsfm_d_inner=2048),
    "e6": dict(channels_input=1, channels_output=1, channels_H=64, max_H=768, encoder_n_layers=6, kernel_size=4,
               stride=2, tsfm_n_layers=3, tsfm_n_head=8, tsfm_d_model=512, tsfm_d_inner=2048),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def synth_noisy(batch, seconds, seed):
    """Cheap synthetic noisy speech-like batch (harmonic stack + coloured noise), (B,1,T) fp32 on the CPU."""
    g = torch.Generator().manual_seed(seed)
    T = int(seconds * SR)
    t = torch.arange(T, dtype=torch.float32) / SR
    f0 = 80 + 220 * torch.rand(batch, 1, generator=g)
    x = torch.zeros(batch, T)
    for k in range(1, 9):
        x += torch.sin(2 * torch.pi * k * f0 * t + 6.28 * torch.rand(batch, 1, generator=g)) / k
    x *= 0.5 * (1 - torch.cos(2 * torch.pi * 4 * t))
    x *= 0.3 / x.abs().amax(1, keepdim=True)
    n = torch.randn(batch, T, generator=g)
    n = 0.5 * (n + torch.roll(n, 1, 1))
    snr = -5 + 30 * torch.rand(batch, 1, generator=g)
    n *= x.pow(2).mean(1, keepdim=True).sqrt() / n.pow(2).mean(1, keepdim=True).sqrt() * 10 ** (-snr / 20)
    return (x + n)[:, None].contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
N_PARAMS_M = {"e8": 41.38, "e6": 27.21}       # README.md:61,121 of the reference (checked by tests/test_cpu_host.py against the constructor)


def workload_name(args):
    """config.workload of the headline bench -- shared by the product arm and the reference arm (same workload, the CPU arm
    times a bounded sample of it per step)."""
    return (f"CleanUMamba {args.model.upper()} full ({N_PARAMS_M[args.model]:.2f}M, seeded random init) offline forward, "
            f"batch {args.batch} x {args.seconds:g} s @16 kHz per GPU")


CPU_SAMPLE_CLIPS = 1      # clips per CPU step (the reference's selective_scan_ref materialises 0.67 GB per 10 s clip at E8)


def shared_config(args, world):
    """`config` of the bench line -- identical for the product arm and the reference arm (same workload; the CPU arm times a
    bounded sample of it per step, described in its cpu_baseline.sample)."""
    return {"workload": workload_name(args), "math": args.math, "global_batch": world * args.batch, "clip_seconds": args.seconds,
            "parallelism": f"utterance-sharded x{world}",
            "l2": "inputs+activations per step (>25 GB) exceed the 126 MB L2; no explicit flush"}


def _reference_net(cfg):
    """The UNMODIFIED reference module (staged copy under baseline/_ref or the /root/reference mount) on the mamba_ssm shim, or
    None when neither is present.  Checker / baseline only: nothing of the product imports this."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_loader
        if not ref_loader.reference_available():
            return None
        refnet = ref_loader.import_reference()
        torch.manual_seed(0)
        return refnet.Net("CleanUMamba", dict(cfg)).float().eval()
    except Exception as exc:      # noqa: BLE001 -- the port remains as the baseline; say why
        print(f"# unmodified reference unavailable: {exc!r}", file=sys.stderr)
        return None


def cpu_reference_pass(cfg, seconds, steps, warmup, threads, clips=CPU_SAMPLE_CLIPS, impl="port"):
    """Times the reference's CPU path on `clips` clips per step; impl = "reference" (unmodified reference files through
    selective_scan_ref) or "port" (oracle/cleanumamba_oracle.py).  Returns (audio_s_per_s, ms_per_step, sample) or None."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    torch.set_num_threads(threads)
    x = synth_noisy(clips, seconds, 1234)
    if impl == "reference":
        net = _reference_net(cfg)
        if net is None:
            return None
        fwd = lambda: net(x.clone())          # noqa: E731  (the reference normalises its input in place)
    else:
        import cleanumamba_oracle as orc
        from cleanumamba_b200.network import Net
        torch.manual_seed(0)
        sd = {k: v.clone() for k, v in Net("CleanUMamba", dict(cfg)).state_dict().items()}
        fwd = lambda: orc.forward(sd, x.clone())          # noqa: E731
    with torch.no_grad():
        for _ in range(warmup):
            fwd()
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd()
        dt = (time.perf_counter() - t0) / steps
    return (clips * seconds / dt, dt * 1e3,
            f"{clips} clip(s) x {seconds:g} s per step (of the batch), {steps} steps after {warmup} warm-up, fp32, {threads} threads")


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)       # same K / W as the product arm; one clip per step keeps it to minutes
    kind, res = "reference", cpu_reference_pass(cfg, args.seconds, steps, warmup, threads, impl="reference")
    port = cpu_reference_pass(cfg, args.seconds, min(steps, 5), 1, threads, impl="port")
    if res is None:
        kind, res = "port", cpu_reference_pass(cfg, args.seconds, steps, warmup, threads, impl="port")
    val, ms, sample = res
    path = ("the reference's own files (src/network/CleanUMamba.py, layers.py, network.py, src/util/util.py: staged unmodified in "
            "baseline/_ref) on CPU through mamba_ssm's selective_scan_ref path (oracle/ref_shim stand-in for the absent wheel), all host cores"
            if kind == "reference" else
            "oracle port of CleanUMamba.forward + mamba_ssm selective_scan_ref (oracle/cleanumamba_oracle.py), all host cores "
            "(baseline/_ref not staged on this box)")
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(args, max(1, args.gpus)),
            "reference_path": path,
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                             "port": {"value": round(port[0], 3), "unit": UNIT, "sample": port[2]} if port else None},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def measure_stream(net, eng, dev, rank, world, dist, S, H, steps, warmup, graph=False, math="f16x3", model="e6", total=None,
                   layout="auto", state_f16=False):
    """configs[2]: S concurrent streams on this GPU, H hops per feed() call, carried conv/SSM state.  A step = one feed() call
    over all streams; value = streamed audio-seconds per wall-second over all ranks.  Returns the record (every rank)."""
    hop = net.total_stride
    g = torch.Generator().manual_seed(4321 + rank)
    n_chunk = H * hop
    host_chunk = (torch.randn(S, n_chunk, generator=g) * 0.1).pin_memory()
    host_out = torch.empty(S, n_chunk).pin_memory()
    chunk = host_chunk.to(dev)
    sess = net.stream_session(batch=S, layout=layout, state_dtype=torch.float16 if state_f16 else torch.float32)
    layout_used = "time_major" if type(sess).__name__ == "TimeMajorStreamSession" else "stream_major"
    sess.feed((torch.randn(S, net.frame_length - hop, generator=g) * 0.1).to(dev))   # prime: next feeds emit H hops each

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        out = sess.feed(chunk)
        assert out.shape == (S, n_chunk), out.shape
    launches_per_replay = 0
    if graph:          # steady-state CUDA graph of feed(): one replay per step instead of ~260 launches + Python glue
        eng.launches = 0
        sess.capture_graph(n_chunk)
        launches_per_replay = eng.launches // 2          # capture_graph runs the step twice (eager rehearsal + capture)
        for _ in range(2):
            sess.feed(chunk)
    barrier()
    eng.prof, eng.launches = (None if graph else []), 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sess.feed(chunk)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches, prof = (launches_per_replay * steps if graph else eng.launches), eng.profile_summary()
    eng.prof = None
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dchunk = torch.empty_like(chunk)
    f0.record()
    for _ in range(steps):
        dchunk.copy_(host_chunk, non_blocking=True)
        host_out.copy_(sess.feed(dchunk), non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    n_streams = S
    if dist is not None:
        from cleanumamba_b200.shard import max_over_ranks
        ms, ms_e2e = max_over_ranks(ms, dev), max_over_ranks(ms_e2e, dev)
        t = torch.tensor([S], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        n_streams = int(t.item())
    audio = n_streams * n_chunk / SR * steps
    kernels = {k: {"ms_per_step": round(v["ms"] / steps, 3), "launches_per_step": v["launches"] // steps}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    del sess
    return {"metric": METRIC + " [streaming]", "value": round(audio / (ms / 1e3), 1), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 3),
            "higher_is_better": True, "scaling": "strong" if total else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"CleanUMamba {model.upper()} streaming, {n_streams} streams in total ({S} on rank 0), {H} hops "
                                   f"({n_chunk} samples, {1e3 * n_chunk / SR:.0f} ms) per feed(), carried conv/SSM state, "
                                   f"math={math}" + (", CUDA-graph replay" if graph else "") +
                                   (", REDUCED-PRECISION variant: SSM state stored as fp16" if state_f16 else ""),
                       "streams_total": n_streams, "streams_per_gpu": S, "buffer_layout": layout_used,
                       "ssm_state_dtype": "f16" if state_f16 else "f32",
                       "real_time_factor_per_stream": round(n_chunk / SR / (ms / steps / 1e3), 2),
                       "chunk_latency_ms": round(ms / steps, 3)},
            "e2e": {"value": round(audio / (ms_e2e / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": S * n_chunk * 4,
                    "d2h_bytes_per_step": S * n_chunk * 4},
            "gpu_launches": launches, "kernels": kernels}


def run_stream(args, net, eng, dev, rank, world, dist):
    from cleanumamba_b200.shard import shard_bounds
    S = args.streams
    if args.streams_total:          # configs[2] as specified: a FIXED total, sharded over the ranks (strong scaling)
        lo, hi = shard_bounds(args.streams_total, rank, world)
        S = hi - lo
    rec = measure_stream(net, eng, dev, rank, world, dist, S, args.hops, args.steps, args.warmup, graph=args.graph, math=args.math,
                         model=args.model, total=args.streams_total, layout=args.layout, state_f16=args.state_f16)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure_train(net, dev, rank, world, dist, B, seconds, steps, warmup, math="f16x3", model="e8", fused_loss=True):
    """configs[3]: E8-full training step (fwd + L1 + multi-resolution STFT loss + bwd + fused Adam), per-GPU batch fixed
    (weak scaling), gradients averaged with the bucketed NCCL all-reduce overlapped with the backward.  At N > 1 the step is
    also timed WITHOUT the all-reduce and the compute stream's stall on NCCL is measured with CUDA events (exposed time)."""
    from cleanumamba_b200.distributed import apply_gradient_allreduce
    from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG, MultiResolutionSTFTLoss, loss_fn
    net.train()
    if dist is not None:
        apply_gradient_allreduce(net)
    opt = torch.optim.Adam(net.parameters(), lr=2e-4, fused=True)
    if fused_loss:      # windowed DFT as a tcgen05 GEMM + fused reductions + own backward (cleanumamba_b200/fused_loss.py)
        from cleanumamba_b200.fused_loss import FusedMultiResolutionSTFTLoss
        mr = FusedMultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)
    else:               # the PyTorch restatement of the reference loss (torch.stft / cuFFT): the checker
        mr = MultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG).to(dev)
    noisy = synth_noisy(B, seconds, 1234 + rank).to(dev)
    clean = synth_noisy(B, seconds, 99 + rank).to(dev) * 0.5
    work = torch.empty_like(noisy)
    eng = net.train_engine()

    def step():
        work.copy_(noisy)
        opt.zero_grad(set_to_none=True)
        loss, _ = loss_fn(net, (clean, work), mrstftloss=mr)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            from cleanumamba_b200.shard import max_over_ranks
            ms = max_over_ranks(ms, dev)
        return ms, loss

    for _ in range(warmup):
        loss = step()
    sync = getattr(net, "_grad_sync", None)
    if sync is not None:
        torch.cuda.synchronize()
        sync.exposed_ms()
    eng.prof, eng.launches = [], 0
    ms, loss = timed(steps)
    launches, prof = eng.launches, eng.profile_summary()
    eng.prof = None
    allreduce = None
    if sync is not None:
        exposed = sync.exposed_ms()
        buckets = list(sync.bucket_bytes)
        net._grad_sync = None                      # same step, no collective: what the all-reduce costs end to end
        ms_local, _ = timed(max(2, steps // 2))
        ms_local /= max(2, steps // 2)
        net._grad_sync = sync
        allreduce = {"bytes": sum(buckets), "buckets": dict(zip(("decoder", "bottleneck", "encoder_deep", "encoder_outer"), buckets)), "overlapped": True,
                     "ms_per_step_without_allreduce": round(ms_local, 3),
                     "exposed_ms_per_step": round(ms / steps - ms_local, 3),
                     "stream_stall_on_nccl_ms_per_step": round(exposed, 3),
                     "limiting_bucket": "encoder_outer (the last one the backward completes: nothing is left to hide it behind; the deep encoder levels "
                                        "-- almost all of the encoder's bytes -- are reduced under the backward of the outer levels)"}
    kernels = {k: {"ms_per_step": round(v["ms"] / steps, 3), "launches_per_step": v["launches"] // steps,
                   "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] and v["ms"] else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    kernel_ms = sum(v["ms"] for v in prof.values()) / steps
    rec = {"metric": METRIC.replace("denoised", "trained on") + " [training step]",
           "value": round(world * B * seconds * steps / (ms / 1e3), 1), "unit": UNIT, "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"CleanUMamba {model.upper()} full training step: fwd + L1 + MR-STFT loss "
                                  f"+ bwd + fused Adam (PyTorch), batch {B} x {seconds:g} s per GPU, math={math}",
                      "loss": "L1 (PyTorch) + fused MR-STFT loss (DFT as tcgen05 GEMM, own backward)" if fused_loss
                              else "L1 + MR-STFT loss in PyTorch (torch.stft / cuFFT)",
                      "global_batch": world * B, "parallelism": f"dp{world}",
                      "grad_allreduce": allreduce,
                      "our_kernels_ms_per_step": round(kernel_ms, 3), "final_loss": round(float(loss), 5)},
           "gpu_launches": launches, "kernels": kernels}
    del opt, mr, noisy, clean, work
    return rec


def run_train(args, net, dev, rank, world, dist):
    B = args.batch if args.batch != 64 else 16
    rec = measure_train(net, dev, rank, world, dist, B, args.seconds, args.steps, args.warmup, math=args.math, model=args.model,
                        fused_loss=args.loss == "fused")
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_sweep(args, net, eng, dev):
    """BASELINE.json configs[4]: long-clip / batch sweep of the offline forward on one GPU -- throughput plus the per-kernel
    times of the scan and of the conv stack (tap-GEMMs) for every (batch, clip length)."""
    points = [(8, 10.0), (64, 10.0), (256, 10.0), (21, 30.0), (85, 30.0), (10, 60.0), (42, 60.0)]
    rows = []

    def tokens(T):
        n = net.valid_length(T)
        for _ in range(net.encoder_n_layers):
            n = (n - 4) // 2 + 1
        return n

    for B, sec in points:
        T = int(sec * SR)
        x_dev = synth_noisy(B, sec, 99).to(dev)
        work = torch.empty_like(x_dev)

        def step():
            work.copy_(x_dev)
            with torch.no_grad():
                return net(work)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        eng.prof, eng.launches = [], 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        prof = eng.profile_summary()
        eng.prof = None
        k = {n: round(v["ms"] / 3, 3) for n, v in prof.items()}
        scan = prof.get("selective_scan")
        rows.append({"batch": B, "clip_seconds": sec, "bottleneck_tokens": tokens(T), "ms_per_step": round(ms, 3),
                     "audio_s_per_s": round(B * sec / (ms / 1e3), 1), "gemm_ms": round(k.get("gemm", 0) + k.get("gemm_tap2", 0), 3),
                     "scan_ms": k.get("selective_scan"),
                     "scan_state_updates_per_s": round(scan["flops"] / (scan["ms"] / 1e3)) if scan else None,
                     "scan_GBps": round(scan["bytes"] / (scan["ms"] / 1e3) / 1e9, 1) if scan else None,
                     "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)})
        print("# sweep", rows[-1], file=sys.stderr, flush=True)
        del x_dev, work
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
    print(json.dumps({"metric": METRIC + " [clip-length / batch sweep]", "unit": UNIT, "n_gpus": 1, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "CleanUMamba E8 full offline forward, math=%s: batch x clip-length sweep (3 timed steps "
                                             "after 2 warm-ups per point, CUDA events)" % args.math},
                      "sweep": rows}))


def run_pruned(args, dev):
    """BASELINE.json configs[0]: the shipped pruned checkpoint E8-pruned-500K (492 K parameters, trained, irregular widths),
    batch 4 x 10 s -- the product on one GPU next to the reference's own CPU path (unmodified reference when staged) on the SAME
    checkpoint and input.  The checkpoint comes from tests/golden/e8_pruned_500k.pt (state_dict + config of the reference's .pkl)."""
    import json as _json
    from cleanumamba_b200.network import Net
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "e8_pruned_500k.pt"), map_location="cpu", weights_only=True)
    cfgp = _json.loads(fx["config"])
    B, sec = (args.batch if args.batch != 64 else 4), args.seconds
    x = synth_noisy(B, sec, 1234)
    net = Net("CleanUMamba", dict(cfgp, math_mode=args.math))
    net.load_pruned_state_dict(fx["state_dict"])
    net = net.to(dev).float().eval()
    host_in, host_out = x.pin_memory(), torch.empty(B, 1, int(sec * SR)).pin_memory()
    x_dev, work = x.to(dev), torch.empty(B, 1, int(sec * SR), device=dev)
    with torch.no_grad():
        for _ in range(max(args.warmup, 5)):
            work.copy_(x_dev); y = net(work)
        torch.cuda.synchronize()
        eng = net.engine()
        eng.launches = 0
        work.copy_(x_dev); eng._forward_eager(work)           # kernels per forward (the timed steps replay them from a CUDA graph)
        per_forward = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            work.copy_(x_dev); y = net(work)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            work.copy_(host_in, non_blocking=True)
            host_out.copy_(net(work), non_blocking=True)
        f1.record()
        torch.cuda.synchronize()
        ms_e2e = f0.elapsed_time(f1) / args.steps
    # CPU arm on the same checkpoint: the unmodified reference when staged, else the oracle port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    kind, fwd = "port", None
    try:
        import ref_loader
        if ref_loader.reference_available():
            refnet = ref_loader.import_reference()
            rn = refnet.Net("CleanUMamba", dict(cfgp))
            rn.load_pruned_state_dict({k: v.clone() for k, v in fx["state_dict"].items()})
            rn = rn.float().eval()
            kind, fwd = "reference", (lambda: rn(x.clone()))
    except Exception as exc:      # noqa: BLE001
        print(f"# unmodified reference unavailable: {exc!r}", file=sys.stderr)
    if fwd is None:
        import cleanumamba_oracle as orc
        fwd = lambda: orc.forward(fx["state_dict"], x.clone())      # noqa: E731
    with torch.no_grad():
        ref_out = fwd()
        t0 = time.perf_counter()
        n_cpu = 5
        for _ in range(n_cpu):
            fwd()
        cpu_ms = (time.perf_counter() - t0) / n_cpu * 1e3
    err = (y.cpu() - ref_out).abs().max().item()
    print(_json.dumps({"metric": METRIC + " [pruned checkpoint]", "value": round(B * sec / (ms / 1e3), 1), "unit": UNIT, "n_gpus": 1,
                       "steps": args.steps, "warmup": max(args.warmup, 5), "ms_per_step": round(ms, 4), "higher_is_better": True,
                       "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                       "config": {"workload": f"CleanUMamba E8 pruned-500K (shipped checkpoint, 491 655 parameters), offline forward, batch {B} x "
                                              f"{sec:g} s @16 kHz, math={args.math}; CUDA-graph replay of the forward", "global_batch": B},
                       "e2e": {"value": round(B * sec / (ms_e2e / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": B * int(sec * SR) * 4,
                               "d2h_bytes_per_step": B * int(sec * SR) * 4, "ms_per_step": round(ms_e2e, 4)},
                       "gpu_launches": per_forward * args.steps,
                       "parity_max_abs_vs_cpu_arm": err,
                       "cpu_baseline": {"value": round(B * sec / (cpu_ms / 1e3), 2), "unit": UNIT, "cores": threads, "kind": kind,
                                        "sample": f"the whole batch ({B} x {sec:g} s) per step, {n_cpu} steps after 1 warm-up, fp32, {threads} threads",
                                        "ms_per_step": round(cpu_ms, 1)}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="e8", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--math", default=os.environ.get("CUM_MATH", "f16x3"), choices=["fp32", "tf32x3", "bf16x3", "f16x3", "bf16", "tf32"],
                    help="arithmetic of the contractions (activations / accumulation / storage are fp32 in every mode): "
                         "f16x3 (default) / tf32x3 = tcgen05 3-pass split products with 22 mantissa bits, parity-tested inside the "
                         "fp32 tolerance of BASELINE.json (max-abs <= 1e-4, dSI-SDR <= 0.01 dB); bf16x3 = 16-17 bits (marginal at "
                         "full-scale amplitude); fp32 = exact CUDA-core FFMA")
    ap.add_argument("--no-variants", action="store_true", help="skip the short tf32x3 / fp32 comparison runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="offline", choices=["offline", "stream", "train", "sweep", "pruned"],
                    help="offline = headline (configs[1]); stream = configs[2]: carried-state chunked inference; "
                         "train = configs[3]: fwd + L1/MR-STFT loss + bwd + Adam, data-parallel gradient all-reduce")
    ap.add_argument("--streams", type=int, default=4096, help="[stream] concurrent streams per GPU")
    ap.add_argument("--hops", type=int, default=16, help="[stream] hops (2^D samples each) per feed() call")
    ap.add_argument("--graph", action="store_true", help="[stream] replay a captured CUDA graph of the steady-state feed()")
    ap.add_argument("--streams-total", type=int, default=0, help="[stream] total streams, sharded over the ranks (overrides --streams)")
    ap.add_argument("--layout", default="auto", choices=["auto", "stream_major", "time_major"], help="[stream] buffer layout of the session")
    ap.add_argument("--state-f16", action="store_true", help="[stream] reduced-precision variant: SSM state stored as fp16 (time-major only)")
    ap.add_argument("--loss", default="fused", choices=["fused", "pytorch"], help="[train] MR-STFT loss implementation")
    ap.add_argument("--no-extras", action="store_true", help="skip the `train` / `stream` sub-records of the default line")
    args = ap.parse_args()
    cfg = CONFIGS[args.model]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback)"
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from cleanumamba_b200.network import Net
    torch.manual_seed(0)
    net = Net("CleanUMamba", dict(cfg, math_mode=args.math)).to(dev).eval()
    eng = net.engine()
    if args.mode == "stream":
        run_stream(args, net, eng, dev, rank, world, dist)
        return
    if args.mode == "train":
        run_train(args, net, dev, rank, world, dist)
        return
    if args.mode == "sweep":
        run_sweep(args, net, eng, dev)
        return
    if args.mode == "pruned":
        del net, eng
        run_pruned(args, dev)
        return
    B, T = args.batch, int(args.seconds * SR)
    host_in = synth_noisy(B, args.seconds, 1234 + rank).pin_memory()
    host_out = torch.empty(B, 1, T).pin_memory()
    x_dev = host_in.to(dev)
    work = torch.empty_like(x_dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        work.copy_(x_dev)              # forward normalises its input in place (reference semantics): fresh copy per step
        with torch.no_grad():
            return net(work)

    from cleanumamba_b200.pipeline import HostPipeline
    pipe = HostPipeline(net)          # public host-buffer API: H2D / forward / D2H of consecutive batches on three streams

    def step_e2e():
        pipe.submit(host_in, host_out)

    for _ in range(args.warmup):
        step_resident()
    barrier()

    # ---- timed region 1: device-resident inputs --------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.prof, eng.launches = [], 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches
    prof = eng.profile_summary()
    eng.prof = None

    # ---- timed region 2: end to end through the public API with host buffers ---------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_e2e()
    pipe.drain()                      # the last D2H copy has landed in host memory
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    # ---- extras (every rank takes part): the collective path and the streaming configuration under the same clock ----------
    extras = {}
    if not args.no_extras and args.model == "e8" and args.math == "f16x3":
        del pipe
        eng._graphs.clear()
        torch.cuda.empty_cache()
        from cleanumamba_b200.shard import shard_bounds
        try:
            torch.manual_seed(0)
            net_t = Net("CleanUMamba", dict(cfg, math_mode=args.math)).to(dev)
            extras["train"] = measure_train(net_t, dev, rank, world, dist, 16, args.seconds, max(3, min(args.steps, 10)), 3, math=args.math)
            del net_t
        except Exception as exc:      # noqa: BLE001 -- the headline line must still be printed; the failure is reported in place
            extras["train"] = {"error": repr(exc)}
        torch.cuda.empty_cache()
        try:
            torch.manual_seed(0)
            net_s = Net("CleanUMamba", dict(CONFIGS["e6"], math_mode=args.math)).to(dev).eval()
            lo, hi = shard_bounds(4096, rank, world)
            # 16 hops per call eagerly; 1 hop per call (4 ms of audio: the latency case) from the captured steady-state CUDA graph
            extras["stream"] = {f"hops{h}": measure_stream(net_s, net_s.engine(), dev, rank, world, dist, hi - lo, h,
                                                            max(3, min(args.steps, 10)), 3, graph=(h == 1), math=args.math, total=4096)
                                for h in (16, 1)}
            # reduced-precision variant, reported separately: the carried SSM state (the HBM floor of a 1-hop call) stored as fp16
            extras["stream"]["hops1_fp16_state_variant"] = measure_stream(net_s, net_s.engine(), dev, rank, world, dist, hi - lo, 1,
                                                                            max(3, min(args.steps, 10)), 3, graph=True, math=args.math,
                                                                            total=4096, state_f16=True)
            # the reference's real-time use -- ONE stream fed hop by hop (module-level feed(): CUDA-graph replay)
            extras["stream"]["single_stream_hop1_graph"] = measure_stream(net_s, net_s.engine(), dev, rank, world, dist, 1, 1, 50, 5,
                                                                            graph=True, math=args.math)
            del net_s
        except Exception as exc:      # noqa: BLE001
            extras["stream"] = {"error": repr(exc)}
        torch.cuda.empty_cache()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    audio_per_step = world * B * args.seconds
    value = audio_per_step * args.steps / (ms / 1e3)
    e2e_value = audio_per_step * args.steps / (ms_e2e / 1e3)
    pk = peaks()

    # roofline of the dominant kernel family (the tap-GEMM: all convs + projections), timed with CUDA events inside
    # the timed region; `peak` is the measured dense bf16 tensor throughput (sustained: timed inside a long step)
    gem = {"ms": 0.0, "launches": 0, "flops": 0}
    for k in ("gemm", "gemm_tap2", "enc0_block", "dec_last_block", "enc_block", "dec_block"):
        if k in prof:
            for f in gem:
                gem[f] += prof[k][f]
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    achieved = gem["flops"] / (gem["ms"] / 1e3) / 1e12 if gem["ms"] else 0.0
    # dram__bytes of the tensor-core launches of one step: from the committed ncu capture of THIS workload and kernel set
    # (profiles/r02_gemm_traffic_f16x3.json, written by tools/launch_table.py from the ncu launch list), null for any other config
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic_f16x3.json")
    if os.path.exists(tpath) and args.math == "f16x3" and B == 64 and args.seconds == 10.0 and args.model == "e8":
        tj = json.load(open(tpath))
        if tj.get("launches_per_step") == gem["launches"] // max(1, args.steps):
            traffic = tj["dram_bytes_per_step_gemm_launches"]
            traffic_note = tj.get("note")
    roofline = {"kernel": "tap-GEMM (conv/convT/1x1/projection contractions, %s)" % args.math, "bound": "tensor",
                "achieved": round(achieved, 2), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / pk["tf_sustained"], 4), "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": pk["src"] + " bf16 sustained",
                "share_of_step": round(gem["ms"] / total_kernel_ms, 4) if total_kernel_ms else None,
                "launches_per_step": gem["launches"] // args.steps,
                "mma_passes": 3 if args.math in ("tf32x3", "bf16x3", "f16x3") else 1,
                "dtype_note": "bf16 storage + bf16 products (reduced precision)" if args.math == "bf16" else None,
                "note": ("achieved = ALGORITHMIC flops (2*M*N*K per contraction) / CUDA-event kernel time; tf32x3 issues 3 "
                         "kind::tf32 MMAs per product (TF32 pipe = 1/2 of the bf16 peak used as denominator), so the pipe-level "
                         "rate is 3x achieved; ncu sm__pipe_tensor_cycles_active = 70-76 % on the K>=1024 layers "
                         "(profiles/r01_ncu_full_gemm_tf32x3.md)") if args.math == "tf32x3" else
                        ("achieved = ALGORITHMIC flops / CUDA-event kernel time; the split modes issue 3 kind::f16 MMAs per product, so the "
                         "tensor-pipe work rate is 3x achieved (mma_work_tflops); per-launch ncu numbers: profiles/r02_launches_f16x3.md")
                        if args.math in ("bf16x3", "f16x3") else None,
                "mma_work_tflops": round(achieved * (3 if args.math in ("tf32x3", "bf16x3", "f16x3") else 1), 1)}
    scan = prof.get("selective_scan")
    scan_roof = None
    if scan:
        gbs = scan["bytes"] / (scan["ms"] / 1e3) / 1e9
        ups = scan["flops"] / (scan["ms"] / 1e3)
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        ceiling = 16 * 148 * mhz * 1e6            # 16 ex2 per clock and SM at the SM clock the box actually held during the timed region
        scan_roof = {"kernel": "selective_scan_fwd", "bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm"],
                     "unit": "GB/s", "frac": round(gbs / pk["hbm"], 4), "traffic": None,
                     "state_updates_per_s": round(ups, 0),
                     "mufu_ceiling_updates_per_s": round(ceiling, 0), "mufu_frac": round(ups / ceiling, 4), "sm_mhz": mhz,
                     "share_of_step": round(scan["ms"] / total_kernel_ms, 4),
                     "note": "d_state=64: one MUFU ex2 per state update -- MUFU-bound, not HBM-bound (SURVEY.md §7); mufu_frac = state "
                             "updates/s over the 16 ex2/clk/SM ceiling at the SM clock sampled during the timed region"}
    kernels = {k: {"ms_per_step": round(v["ms"] / args.steps, 3), "launches_per_step": v["launches"] // args.steps}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    variants = None
    if world == 1 and not args.no_variants and args.mode == "offline":
        # the other arithmetic modes on the same workload (2 timed steps each after 1 warm-up), for transparency
        variants = {}
        for alt in ("f16x3+fp16-stored-weights", "bf16", "tf32x3", "bf16x3", "fp32"):
            if alt == args.math:
                continue
            torch.manual_seed(0)
            net_alt = Net("CleanUMamba", dict(cfg, math_mode=alt.split("+")[0])).to(dev).eval()
            net_alt.load_state_dict(net.state_dict())
            if "+" in alt:      # weights rounded to fp16 exactly as the reference's released checkpoints store them:
                with torch.no_grad():       # the low weight halves vanish and the engine runs 2 MMA passes per product
                    for p_ in net_alt.parameters():
                        p_.copy_(p_.half().float())

            def alt_step():
                work.copy_(x_dev)
                with torch.no_grad():
                    return net_alt(work)
            alt_step()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                alt_step()
            a1.record()
            torch.cuda.synchronize()
            variants[alt] = {"value": round(B * args.seconds * 2 / (a0.elapsed_time(a1) / 1e3), 1), "unit": UNIT,
                             "ms_per_step": round(a0.elapsed_time(a1) / 2, 3)}
            del net_alt
            torch.cuda.empty_cache()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        # ~10-20 s of CPU work: the unmodified reference (staged in baseline/_ref) when present, the oracle port beside it
        ref = cpu_reference_pass(cfg, args.seconds, 5, 1, threads, impl="reference")
        port = cpu_reference_pass(cfg, args.seconds, 8, 1, threads, impl="port")
        v, _, sample = ref if ref else port
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "reference" if ref else "port", "sample": sample,
               "port": {"value": round(port[0], 3), "unit": UNIT, "sample": port[2]}}

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "arithmetic": {"fp32": "fp32 storage + exact fp32 FFMA products",
                           "tf32x3": "fp32 storage/accumulate; products = 3 TF32 tensor-core passes on hi/lo halves (~2^-21 per product)",
                           "bf16x3": "fp32 storage/accumulate; products = 3 bf16 tensor-core passes on hi/lo halves (~2^-16 per product); "
                                     "parity-tested at this workload size against the exact-fp32 mode: max-abs <= 1e-4, dSI-SDR <= 0.01 dB",
                           "f16x3": "fp32 storage/accumulate; products = 3 fp16 tensor-core passes on hi/lo halves (11+11 bits, ~2^-21 per "
                                    "product, same accuracy class as tf32x3 at the bf16 tensor rate; weights carry a power-of-two scale, "
                                    "activations convert with saturation); parity-tested at this workload size against the exact-fp32 mode",
                           "bf16": "REDUCED PRECISION variant (reported separately, outside the fp32 tolerance): bf16 activation storage and "
                                   "single-pass bf16 tensor-core products in the encoder / decoder stacks, fp32 accumulate",
                           "tf32": "single TF32 pass (outside the tolerance)"}[args.math],
            "config": shared_config(args, world),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": B * T * 4,
                    "d2h_bytes_per_step": B * T * 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "scan_roofline": scan_roof,
            "kernels": kernels}
    if variants:
        line["variants"] = variants
    if cpu:
        line["cpu_baseline"] = cpu
    line.update(extras)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
