#!/usr/bin/env python
"""Headline benchmark: UiT 1 s-clip inferences/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config headline|c2|c3|c4|c5] [--arch ..] [--batch ..]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path over one batch of synthetic input.  `--config` picks the BASELINE.json workload:
  headline  UiT-XS, 4096 synthetic 1 s clips PER GPU (weak scaling; the line the driver records)          [default]
  c2        UiT-XXXS, 4096 x 1 s clips on one B200                                                        (configs[1])
  c3        UiT-XXS, 65 536 x 1 s clips in total, clips sharded over the N GPUs (strong scaling)          (configs[2])
  c4        log-mel front-end alone, 16 384 x 10 s clips in total, sharded (strong; HBM roofline sweep)   (configs[3])
  c5        UiT-XS sliding 1 s windows (hop 1600) over one 57.6 M-sample stream; every rank takes a time range plus a
            14 400-sample halo, ONE top-dB scope over all windows (cross-rank max)              (strong)   (configs[4])
`value` is measured with the inputs resident in HBM, `e2e` through the host-buffer entry point (pinned host buffers in,
host scores out, copies inside the timed region).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# Algorithmic work per 1 s clip: encoder FLOPs (2*M*N*K over patch embed, block
# GEMMs, both attention matmuls, head) and front-end bytes (4*L + 4*64*T).
ENCODER_FLOPS = {"uit_xs": 68.66e6, "uit_xxs": 35.18e6, "uit_xxxs": 24.03e6}
LOGMEL_BYTES_1S = 4 * 16000 + 4 * 64 * 101
LOGMEL_BYTES_10S = 4 * 160000 + 4 * 64 * 1001
CPU_SAMPLE_CLIPS = 512
CONFIGS = {
    "headline": dict(arch="uit_xs", per_gpu=4096, scaling="weak", what="1 s clips"),
    "c2": dict(arch="uit_xxxs", per_gpu=4096, scaling="weak", what="1 s clips"),
    "c3": dict(arch="uit_xxs", total=65536, scaling="strong", what="1 s clips"),
    "c4": dict(arch="uit_xs", total=16384, scaling="strong", what="10 s clips, front-end only"),
    "c5": dict(arch="uit_xs", stream=57_600_000, hop=1600, scaling="strong", what="sliding 1 s windows"),
}


class stdout_to_stderr:
    """NCCL prints its version banner on the C-level stdout when the first communicator is created; the contract is ONE
    JSON line on stdout, so fd 1 is pointed at stderr while the process group is set up and warmed up."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def tensor_peak(peaks, timed_seconds: float):
    """The burst figure for a timed region under ~1 s (isolation conditions: clocks at max), the sustained one inside a
    long step (the recipe's rule)."""
    if timed_seconds < 1.0:
        return peaks["bf16_tflops"], peaks["source"] + " (bf16 cuBLAS, burst: timed region < 1 s)"
    return peaks["bf16_tflops_sustained"], peaks["source"] + " (bf16 cuBLAS, sustained: timed region >= 1 s)"


def ncu_traffic(kernel: str, arch: str, batch: int):
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/traffic.json), if it was taken on this
    configuration; else None.  NOT measured in this run (ncu cannot run inside the timed bench)."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(p) or arch != "uit_xs" or batch != 4096:
        return None
    with open(p) as f:
        d = json.load(f)
    return d.get(kernel, {}).get("traffic_bytes")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + ",".join(self.FIELDS),
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) != len(self.FIELDS):
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (same ATen CPU ops in the same order as the reference; pinned against it by
# tests/golden/generate_golden.py).  ONE method for the in-arm cpu_baseline and for `--impl reference`.
# ------------------------------------------------------------------------------------------------------------------------
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_workload(config: str, arch: str):
    """(state_dict, fn, units per pass, description) of the bounded CPU sample of `config`."""
    import torch
    import uit_mobile_b200 as U
    from oracle import uit_oracle as O
    torch.manual_seed(0)
    model = getattr(U.models, arch)(outputdim=537, target_length=102)      # parameter holders on CPU (no kernel involved)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    if config == "c4":
        x = (0.1 * torch.randn(64, 160000, generator=g)).clamp_(-1, 1)
        win, fb = sd["front_end.0.spectrogram.window"], sd["front_end.0.mel_scale.fb"]
        return sd, (lambda: O.logmel(x, win, fb)), 64, "64 synthetic 10 s clips per pass, front-end only"
    if config == "c5":
        stream = (0.1 * torch.randn(16000 + 1600 * (CPU_SAMPLE_CLIPS - 1), generator=g)).clamp_(-1, 1)
        xw = stream.unfold(0, 16000, 1600).contiguous()
        return sd, (lambda: O.forward(sd, xw)), CPU_SAMPLE_CLIPS, f"{CPU_SAMPLE_CLIPS} sliding 1 s windows (hop 1600) per pass"
    x = (0.1 * torch.randn(CPU_SAMPLE_CLIPS, 16000, generator=g)).clamp_(-1, 1)
    return sd, (lambda: O.forward(sd, x)), CPU_SAMPLE_CLIPS, f"{CPU_SAMPLE_CLIPS} synthetic 1 s clips per pass"


def time_cpu(fn, units: int, threads: int, warmup: int, steps: int = 0, min_seconds: float = 0.0, max_steps: int = 64):
    """`warmup` untimed passes, then `steps` timed passes (or as many as fit `min_seconds`, at least 2): units/s over the
    whole timed region."""
    import torch
    torch.set_num_threads(threads)
    for _ in range(warmup):
        fn()
    n, t0 = 0, time.perf_counter()
    while (n < steps) if steps else (n < 2 or (time.perf_counter() - t0 < min_seconds and n < max_steps)):
        fn()
        n += 1
    dt = time.perf_counter() - t0
    return units * n / dt, n, dt


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is pure Python and
    /root/reference does not exist on the GPU box, so this times the oracle port with ALL host threads (torchrun sets
    OMP_NUM_THREADS=1: the thread count is set explicitly), on a bounded sample per step.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = CONFIGS[args.config]
    arch = args.arch or cfg["arch"]
    threads = host_threads()
    sd, fn, units, what = cpu_workload(args.config, arch)
    value, n, dt = time_cpu(fn, units, threads, warmup=args.warmup, steps=args.steps)
    sample = f"{what} (bounded sample of the workload), fp32, {threads} threads"
    line = {
        "impl": "reference", "metric": metric_name(args.config, arch), "value": value, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * dt / n,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, arch, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(config: str, arch: str) -> str:
    if config == "c4":
        return "log-mel front-end 10s-clip inferences/sec"
    if config == "c5":
        return f"{arch} sliding 1s-window inferences/sec"
    return f"{arch} 1s-clip inferences/sec"


def workload_config(args, arch: str, world: int):
    cfg = CONFIGS[args.config]
    if "per_gpu" in cfg:
        B = args.batch or cfg["per_gpu"]
        return {"workload": f"{arch} batched inference, {B} synthetic 1 s clips (16 kHz) per GPU, random-init weights",
                "baseline_config": args.config, "arch": arch, "batch_per_gpu": B, "global_batch": B * world, "clip_samples": 16000,
                "precision": args.precision, "parallelism": f"clips sharded x{world}" if world > 1 else "single GPU",
                "l2": f"inputs ({B * 64000 / 1e6:.0f} MB/step) larger than the 126 MB L2"}
    if args.config == "c3":
        total = args.batch or cfg["total"]
        return {"workload": f"{arch} batched inference, {total} synthetic 1 s clips in total, random-init weights (BASELINE configs[2])",
                "baseline_config": "c3", "arch": arch, "global_batch": total, "clip_samples": 16000, "precision": args.precision,
                "parallelism": f"clips sharded x{world} (tile-aligned contiguous slices)", "l2": "inputs larger than the 126 MB L2"}
    if args.config == "c4":
        total = args.batch or cfg["total"]
        return {"workload": f"log-mel front-end alone, {total} synthetic 10 s clips in total (BASELINE configs[3])",
                "baseline_config": "c4", "global_batch": total, "clip_samples": 160000, "parallelism": f"clips sharded x{world}",
                "l2": "inputs larger than the 126 MB L2"}
    n = cfg["stream"]
    return {"workload": f"{arch} sliding 1 s windows, hop {cfg['hop']}, over one {n}-sample synthetic stream (1 h; BASELINE configs[4])",
            "baseline_config": "c5", "arch": arch, "stream_samples": n, "hop": cfg["hop"], "windows": (n - 16000) // cfg["hop"] + 1,
            "precision": args.precision, "parallelism": f"time ranges + 14400-sample halo x{world}, one top-dB scope (cross-rank max)",
            "l2": "stream slice larger than the 126 MB L2 at N <= 1" if world == 1 else "per-rank slice + log-mel exceed L2 only at N <= 2"}


# ------------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import uit_mobile_b200 as U
    from uit_mobile_b200 import _native as N
    from uit_mobile_b200 import sharding
    from uit_mobile_b200.pipeline import FrontEndHostPipeline, HostPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nnodes=1 "
                             f"--nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...")
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the UiT hot path has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    quiet = stdout_to_stderr()
    quiet.__enter__()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = N.lib()
    peaks = measured_peaks()
    cfg = CONFIGS[args.config]
    arch = args.arch or cfg["arch"]
    steps = args.steps if args.steps else {"headline": 300, "c2": 300, "c3": 30, "c4": 10, "c5": 30}[args.config]
    warmup = max(args.warmup, 3)

    torch.manual_seed(0)                                        # identical random-init weights on every rank
    model = getattr(U.models, arch)(outputdim=537, target_length=102, precision=args.precision).to(dev).eval()
    max_exchange = None
    if world > 1 and args.diag != "nogroup":
        model.process_group = dist.group.WORLD
        max_exchange = "nccl all_reduce(MAX) of one word, async under the speculative encode"
        if not args.nccl_gather:
            why = ""
            try:
                model.peer_words = sharding.PeerWords(dist.group.WORLD, dev)
            except Exception as e:                    # no peer mapping on this box: NCCL
                why = f"{type(e).__name__}: {str(e)[:100]}"
            ok = torch.tensor([0 if model.peer_words is None else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)            # every rank takes the same path, or none does
            if int(ok.item()):
                max_exchange = "PeerWords: publish / collect kernels over NVLink peer memory (no NCCL call in the step)"
            else:
                model.peer_words = None
                max_exchange += f" (PeerWords unavailable on some rank: {why})"
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    pending = [None]      # in-flight all-gather of the previous step (NCCL stream): overlaps the next step's kernels

    peer = [None, "nccl all_gather_into_tensor"]      # PeerGather (copy engines over NVLink peer memory) once the shard sizes are known

    def gather_async(probs, total, sizes_equal=True):
        """all-gather of the scores, left in flight: it is waited for at the start of the NEXT step's gather (and drained inside
        the timed region after the last step).  Copy-engine pulls over NVLink peer memory when the group can map it, else NCCL."""
        if pending[0] is not None:
            pending[0][1].wait()
        if peer[0] is not None:
            out, done = peer[0](probs)

            class _W:
                def wait(self_inner):
                    torch.cuda.current_stream().wait_event(done)
            pending[0] = (out, _W(), probs)
            return out
        out = torch.empty((total, probs.shape[1]), dtype=probs.dtype, device=dev)
        pending[0] = (out, dist.all_gather_into_tensor(out, probs, async_op=True), probs)
        return out

    def drain():
        if pending[0] is not None:
            pending[0][1].wait()
            pending[0] = None

    # ---- workload -----------------------------------------------------------------------------------------------------
    kind = "clips"
    L = 16000
    if "per_gpu" in cfg:
        B = args.batch or cfg["per_gpu"]
        total = B * world
    elif args.config in ("c3", "c4"):
        total = args.batch or cfg["total"]
        L = 160000 if args.config == "c4" else 16000
        align = model.tile_clips(1 + L // 160) if args.config == "c3" else 1
        b0, b1 = sharding.shard_bounds(total, rank, world, align)
        B = b1 - b0
        if args.config == "c4":
            kind = "frontend"
    else:
        kind = "sliding"
        n_stream, hop = cfg["stream"], cfg["hop"]
        total = (n_stream - 16000) // hop + 1
        w0, w1, s0, s1 = sharding.window_shard_bounds(total, hop, 16000, rank, world)
        B = w1 - w0
    if kind == "sliding":
        x = (0.1 * torch.randn(s1 - s0, generator=g, device=dev)).clamp_(-1, 1)          # this rank's time range + halo
    else:
        x = torch.empty((B, L), dtype=torch.float32, device=dev)
        for i in range(0, B, 2048):                                                      # chunked: randn temporaries stay small
            x[i:i + 2048] = (0.1 * torch.randn(min(2048, B - i), L, generator=g, device=dev)).clamp_(-1, 1)
    if kind == "sliding":
        shard_sizes = [sharding.shard_bounds(total, r, world)[1] - sharding.shard_bounds(total, r, world)[0] for r in range(world)]
        g_align = 1
    elif "per_gpu" in cfg:
        shard_sizes, g_align = [B] * world, 1
    else:
        shard_sizes = [sharding.shard_bounds(total, r, world, align)[1] - sharding.shard_bounds(total, r, world, align)[0] for r in range(world)]
        g_align = align
    equal_shards = len(set(shard_sizes)) == 1
    last_local = [None]
    if world > 1 and kind != "frontend" and not args.nccl_gather:
        why = ""
        try:
            peer[0] = sharding.PeerGather(shard_sizes, 537, dist.group.WORLD, dev, depth=2)
        except Exception as e:                        # no peer mapping on this box: NCCL
            why = f"{type(e).__name__}: {str(e)[:120]}"
        ok = torch.tensor([0 if peer[0] is None else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                # every rank takes the same path, or none does
        if int(ok.item()):
            peer[1] = "PeerGather: copy-engine pulls over NVLink peer memory (torch symmetric memory), no SMs"
            equal_shards = True                       # PeerGather takes ragged shards as they are
        else:
            peer[0] = None
            peer[1] = f"nccl all_gather (PeerGather unavailable on some rank: {why})"

    def step(marks=None):
        with torch.no_grad():
            if kind == "frontend":
                db = model.front_end(x)                         # K1 + the (device-conditional) top-dB clamp pass
                if marks is not None:
                    marks[1].record(); marks[2].record()
                return db
            words = model._new_words(dev)
            if kind == "sliding":
                db, _ = model.front_end.logmel_sliding(x, 16000, hop, max_pow=words[0:1], min_pow=words[1:2])
            else:
                db, _ = model.front_end.logmel_unclamped(x, max_pow=words[0:1], min_pow=words[1:2])
            if marks is not None:
                marks[1].record()
            probs = model._finish(db, words)       # encoder (+ async max all-reduce and conditional exact re-run when sharded)
            if marks is not None:
                marks[2].record()
            last_local[0] = probs
            if world > 1 and equal_shards:
                probs = gather_async(probs, total)
            elif world > 1:
                probs = sharding.gather_scores(probs, total, align=g_align)
        return probs

    for _ in range(warmup):
        step()
    drain()
    barrier()
    quiet.__exit__()
    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- pass 1, launches serialised on one stream: the per-kernel durations behind the roofline entries --------------------------
    pipelined = kind == "clips" and not args.no_pipeline
    seq_steps = min(steps, 50) if pipelined else steps
    launches0 = lib.uitk_kernel_launches()
    marks = [[ev(), ev(), ev()] for _ in range(seq_steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(seq_steps):
        marks[i][0].record()
        probs = step(marks[i])
    host_issue_ms = 1e3 * (time.perf_counter() - t_host0) / seq_steps  # CPU time to ISSUE a step (launch-bound if >= ms_per_step)
    drain()                 # the last all-gather completes inside the timed region
    e1.record()
    barrier()
    launches = lib.uitk_kernel_launches() - launches0
    ms_seq = max_over_ranks(e0.elapsed_time(e1))
    ms_logmel = statistics.mean(m[0].elapsed_time(m[1]) for m in marks)
    ms_encoder = statistics.mean(m[1].elapsed_time(m[2]) for m in marks)
    sequential = {"value": total * seq_steps / (ms_seq * 1e-3), "ms_per_step": ms_seq / seq_steps, "steps": seq_steps}
    ms_total, timed_steps = ms_seq, seq_steps

    # ---- pass 2 (1 s-clip batches), the headline: the same steps through BatchPipeline, two batches in flight - the front-end of
    # step i+1 on one stream under the encoder tail of step i on another; same kernels, same order per batch, same bits
    if pipelined:
        from uit_mobile_b200.pipeline import BatchPipeline
        bp = BatchPipeline(model, depth=3 if world > 1 else 2)

        def pipelined_steps(n):
            prev = None
            for _ in range(n):
                t = bp.submit(x)
                if prev is not None:
                    y = bp.result(prev)
                    last_local[0] = y
                    if world > 1 and args.diag != "nogather":
                        gather_async(y, total) if equal_shards else sharding.gather_scores(y, total, align=g_align)
                prev = t
            y = bp.result(prev)
            last_local[0] = y
            if world > 1 and args.diag != "nogather":
                gather_async(y, total) if equal_shards else sharding.gather_scores(y, total, align=g_align)
            drain()

        pipelined_steps(warmup)
        barrier()
        launches0 = lib.uitk_kernel_launches()
        p0, p1 = ev(), ev()
        p0.record()
        pipelined_steps(steps)
        p1.record()
        barrier()
        launches = lib.uitk_kernel_launches() - launches0
        ms_total, timed_steps = max_over_ranks(p0.elapsed_time(p1)), steps
        seq_probs = probs if world == 1 else None
        if seq_probs is not None and not torch.equal(seq_probs, last_local[0]):
            raise SystemExit("BatchPipeline result differs from the serialised path")
    clocks = sampler.stop() if sampler else None
    value = total * timed_steps / (ms_total * 1e-3)
    steps = timed_steps

    # ---- end to end through the host-buffer entry point (pinned host in, host results out) ---------------------------------
    e2e_steps = max(3, min(steps, 30 if kind == "clips" else 5))
    x_host = x.cpu().pin_memory()
    h2d_floor = None
    e2e_pipelined = False
    if kind == "clips":
        pipe = HostPipeline(model, B, 16000, chunk=args.chunk)
        run_e2e = lambda: pipe(x_host)
        e2e_pipelined = True           # timed through submit() / result(): two host batches in flight
    elif kind == "frontend":
        x = None
        torch.cuda.empty_cache()
        pipe = FrontEndHostPipeline(model, B, L, chunk=256)
        run_e2e = lambda: pipe(x_host)
    else:
        scores_host = torch.empty((B, 537), dtype=torch.float32).pin_memory()

        class _P:
            h2d_bytes = x_host.numel() * 4
            d2h_bytes = B * 537 * 4
            chunk = None
        pipe = _P()

        def run_e2e():
            with torch.no_grad():
                xs = x_host.to(dev, non_blocking=True)
                scores_host.copy_(model.forward_sliding(xs, hop=hop), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return scores_host
    def e2e_loop(p, src, n):
        """n batches through the host pipeline, two in flight: every batch is uploaded from pinned host memory and its scores
        are read back to the host inside the loop (result() blocks until the batch's last download has landed)."""
        prev, out = None, None
        for _ in range(n):
            t = p.submit(src)
            if prev is not None:
                out = p.result(prev)
            prev = t
        return p.result(prev)

    for _ in range(2):
        out_host = run_e2e()
    barrier()
    e2, e3 = ev(), ev()
    e2.record()
    if e2e_pipelined:
        out_host = e2e_loop(pipe, x_host, e2e_steps)
    else:
        for _ in range(e2e_steps):
            out_host = run_e2e()
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    e2e_sync = None
    if e2e_pipelined:                 # the one-batch-at-a-time call for comparison (the tail of every batch is exposed)
        barrier()
        e8, e9 = ev(), ev()
        e8.record()
        for _ in range(e2e_steps):
            out_host = run_e2e()
        e9.record()
        barrier()
        ms_sync = max_over_ranks(e8.elapsed_time(e9))
        e2e_sync = {"value": total * e2e_steps / (ms_sync * 1e-3), "ms_per_step": ms_sync / e2e_steps,
                    "note": "pipe(x_host): one batch at a time, host-synchronous"}
    e2e_value = total * e2e_steps / (ms_e2e * 1e-3)
    e2e_ok = bool(torch.equal(out_host.to(dev), last_local[0])) if kind in ("clips", "sliding") else None
    # the host link's floor: bare pinned H2D of the same bytes on every rank at once (explains e2e scaling: all GPUs of the
    # box share the host's memory / PCIe root complexes)
    stage = torch.empty(min(x_host.numel(), 64 << 20), dtype=torch.float32, device=dev)
    src = x_host.view(-1)[:stage.numel()]
    for _ in range(2):
        stage.copy_(src, non_blocking=True)
    barrier()
    e4, e5 = ev(), ev()
    e4.record()
    for _ in range(5):
        stage.copy_(src, non_blocking=True)
    e5.record()
    barrier()
    h2d_floor = 5 * stage.numel() * 4 / (max_over_ranks(e4.elapsed_time(e5)) * 1e-3) / 1e9
    del stage

    e2e16 = None
    if kind == "clips" and args.config == "headline":
        # the same through 16-bit PCM host buffers (separate, labelled mode: SURVEY 8f n3; halves the H2D bytes)
        pcm_host = (x_host * 32767.0).round().to(torch.int16).pin_memory()
        pipe16 = HostPipeline(model, B, 16000, chunk=args.chunk, dtype=torch.int16)
        for _ in range(2):
            pipe16(pcm_host)
        barrier()
        e6, e7 = ev(), ev()
        e6.record()
        e2e_loop(pipe16, pcm_host, e2e_steps)
        e7.record()
        barrier()
        ms16 = max_over_ranks(e6.elapsed_time(e7))
        e2e16 = {"value": total * e2e_steps / (ms16 * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": pipe16.h2d_bytes,
                 "d2h_bytes_per_step": pipe16.d2h_bytes, "ms_per_step": ms16 / e2e_steps,
                 "note": "same path fed 16-bit PCM host buffers (x = pcm/32768 in-kernel); not the headline"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    timed_s = ms_total * 1e-3
    tpeak, tsrc = tensor_peak(peaks, timed_s)
    line = {
        "metric": metric_name(args.config, arch), "value": value, "unit": "clips/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32" if (kind == "frontend" or args.precision == "fp32") else "bf16",
        "data": "synthetic", "config": workload_config(args, arch, world), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "chunk": pipe.chunk, "matches_device_path": e2e_ok,
                "api": ("HostPipeline.submit / result, two host batches in flight (every batch: H2D from pinned host memory, "
                        "scores D2H and read on the host)") if e2e_pipelined else "host-synchronous call per step",
                "host_synchronous": e2e_sync,
                "h2d_floor_gbs_per_gpu": h2d_floor,
                "h2d_floor_note": f"bare pinned-host->device copy, all {world} rank(s) at once; the e2e step moves "
                                  f"{pipe.h2d_bytes / 1e6:.0f} MB in = {pipe.h2d_bytes / 1e6 / max(h2d_floor, 1e-9):.2f} ms at that rate"},
        "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms,
        "timed_region": ("BatchPipeline: two batches in flight (front-end of step i+1 under the encoder tail of step i)" if pipelined
                         else "launches serialised on one stream"),
        "sequential": sequential, "score_gather": peer[1] if world > 1 else None, "max_word_exchange": max_exchange, **({"diag_INVALID_AS_BENCH": args.diag} if args.diag else {}),
    }
    if e2e16:
        line["e2e_int16_pcm"] = e2e16
    units = B                                                   # per-GPU units behind the per-launch kernel times
    if kind == "frontend":
        fe_gbs = units * LOGMEL_BYTES_10S / (ms_logmel * 1e-3) / 1e9
        line["roofline"] = {"kernel": "logmel_kernel (+ conditional clamp pass)", "bound": "hbm", "achieved": fe_gbs, "peak": peaks["hbm_gbs"],
                            "unit": "GB/s", "frac": fe_gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                            "ms_per_launch": ms_logmel, "bytes_per_clip": LOGMEL_BYTES_10S, "clips_per_launch": units}
    else:
        flops = ENCODER_FLOPS[arch]
        enc_tflops = units * flops / (ms_encoder * 1e-3) / 1e12
        line["roofline"] = {"kernel": "encoder (uitk_encoder: encoder_tc_kernel + head)", "bound": "tensor", "achieved": enc_tflops,
                            "peak": tpeak, "unit": "TFLOP/s", "frac": enc_tflops / tpeak,
                            "frac_vs_sustained": enc_tflops / peaks["bf16_tflops_sustained"],
                            "traffic": ncu_traffic("encoder_tc_kernel", arch, units) if args.precision == "bf16" else None,
                            "traffic_unit": "bytes/launch (dram read+write)",
                            "traffic_source": "committed ncu --set full capture (profiles/traffic.json), not measured in this run",
                            "peak_source": tsrc, "ms_per_launch": ms_encoder,
                            "ms_per_launch_source": "CUDA events around the launch in the serialised pass" + (
                                "; at N > 1 the span also holds the wait for the max-word all-reduce of the slowest rank (ranks are not re-aligned "
                                "between steps): quote the kernel's roofline from the N = 1 line" if world > 1 else ""),
                            "flops_per_clip": flops, "clips_per_launch": units}
        if kind == "clips":
            fe_gbs = units * LOGMEL_BYTES_1S / (ms_logmel * 1e-3) / 1e9
            line["roofline_frontend"] = {"kernel": "logmel_kernel", "bound": "hbm", "achieved": fe_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                         "frac": fe_gbs / peaks["hbm_gbs"], "traffic": ncu_traffic("logmel_kernel", arch, units),
                                         "traffic_unit": "bytes/launch (dram read+write)",
                                         "traffic_source": "committed ncu --set full capture (profiles/traffic.json), not measured in this run",
                                         "peak_source": peaks["source"], "ms_per_launch": ms_logmel, "bytes_per_clip": LOGMEL_BYTES_1S}
        else:
            line["frontend_ms_per_launch"] = ms_logmel

    if args.config == "headline" and world == 1 and not args.no_extras:
        with torch.no_grad():
            # BASELINE config 4 shape (front-end alone on 10 s clips), bounded to 1024 clips so that the default run stays short
            # (the full 16 384-clip run is `--config c4`)
            n10 = 1024
            x10 = (0.1 * torch.randn(n10, 160000, generator=g, device=dev)).clamp_(-1, 1)       # 655 MB > L2
            for _ in range(3):
                model.front_end.logmel_unclamped(x10)
            torch.cuda.synchronize()
            f0, f1 = ev(), ev()
            f0.record()
            for _ in range(5):
                model.front_end.logmel_unclamped(x10)
            f1.record()
            torch.cuda.synchronize()
            ms10 = f0.elapsed_time(f1) / 5
            gbs10 = n10 * LOGMEL_BYTES_10S / (ms10 * 1e-3) / 1e9
            line["frontend_10s"] = {"workload": f"log-mel front-end alone, {n10} synthetic 10 s clips (BASELINE config 4 shape, bounded; full size: --config c4)",
                                    "clips_per_s": n10 / (ms10 * 1e-3), "achieved": gbs10, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                    "frac": gbs10 / peaks["hbm_gbs"], "bytes_per_clip": LOGMEL_BYTES_10S, "ms_per_launch": ms10}
            del x10
            # BASELINE config 5 shape, one GPU's share (57.6 M samples / 8 GPUs); the full stream is `--config c5`
            n_s, hop_s = 7_200_000, 1600
            stream = (0.1 * torch.randn(n_s, generator=g, device=dev)).clamp_(-1, 1)
            n_win = (n_s - 16000) // hop_s + 1

            def per_window():       # every window runs the whole front-end (windows read in place, row stride = hop)
                db_w, mp_w = model.front_end.logmel_unclamped(stream, ld=hop_s, B=n_win, L=16000)
                return model.encode(db_w, mp_w)

            ms_sl = {}
            for name, fn in (("per_window_frontend", per_window), ("shared_stft", lambda: model.forward_sliding(stream, hop=hop_s))):
                for _ in range(2):
                    y_sl = fn()
                torch.cuda.synchronize()
                f0, f1 = ev(), ev()
                f0.record()
                for _ in range(5):
                    y_sl = fn()
                f1.record()
                torch.cuda.synchronize()
                ms_sl[name] = (f0.elapsed_time(f1) / 5, y_sl)
            line["sliding_windows"] = {
                "workload": f"{n_win} sliding 1 s windows, hop {hop_s}, over one {n_s}-sample stream (BASELINE config 5 shape, one GPU's share)",
                "windows_per_s": n_win / (ms_sl["shared_stft"][0] * 1e-3), "ms": ms_sl["shared_stft"][0],
                "ms_per_window_frontend": ms_sl["per_window_frontend"][0],
                "bit_identical_to_per_window_path": bool(torch.equal(ms_sl["shared_stft"][1], ms_sl["per_window_frontend"][1]))}
            del stream
    if world == 1 and not args.no_cpu_baseline:
        from oracle import uit_oracle as O
        threads = host_threads()
        sd, fn, cunits, what = cpu_workload(args.config, arch)
        cps, n, dt = time_cpu(fn, cunits, threads, warmup=1, min_seconds=10.0, max_steps=16)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                                "sample": f"{what}, {n} timed passes after 1 warm-up ({dt:.1f} s), all host threads"}
        if args.config == "headline" and not args.no_extras:
            # the second CPU row (the reference trainer's default: 1 thread, run.py:66-70) ...
            import torch as _t
            sd1 = sd
            x1 = (0.1 * _t.randn(64, 16000, generator=_t.Generator().manual_seed(5))).clamp_(-1, 1)
            cps1, n1, dt1 = time_cpu(lambda: O.forward(sd1, x1), 64, 1, warmup=1, min_seconds=4.0, max_steps=8)
            line["cpu_baseline_1thread"] = {"value": cps1, "unit": "clips/s", "cores": 1, "kind": "port",
                                            "sample": f"64 synthetic 1 s clips per pass, {n1} passes ({dt1:.1f} s), torch.set_num_threads(1)"}
            _t.set_num_threads(threads)
            # ... and the informational second comparator: the same reference path as eager PyTorch ON THE B200 (cuFFT via
            # torch.stft, cuBLAS, ATen elementwise: the oracle port run with its tensors on `cuda`) - the strongest same-box
            # implementation of the reference path, and the practical bar for the front-end kernel
            sd_gpu = {k: v.to(dev) for k, v in sd.items()}
            xg = (0.1 * _t.randn(4096, 16000, generator=g, device=dev)).clamp_(-1, 1)
            win_g, fb_g = sd_gpu["front_end.0.spectrogram.window"], sd_gpu["front_end.0.mel_scale.fb"]

            def dev_ms(f, iters):
                for _ in range(2):
                    f()
                torch.cuda.synchronize()
                a, b = ev(), ev()
                a.record()
                for _ in range(iters):
                    f()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / iters
            ms_eager_fe = dev_ms(lambda: O.logmel(xg, win_g, fb_g), 10)
            ms_eager = dev_ms(lambda: O.forward(sd_gpu, xg), 5)
            line["gpu_eager_comparator"] = {
                "what": "reference path as eager PyTorch fp32 on the same B200 (torch.stft -> cuFFT, cuBLAS, ATen), 4096 x 1 s clips, device-resident",
                "clips_per_s": 4096 / (ms_eager * 1e-3), "ms_per_step": ms_eager,
                "frontend_only_ms": ms_eager_fe, "frontend_only_gbs": 4096 * LOGMEL_BYTES_1S / (ms_eager_fe * 1e-3) / 1e9,
                "ours_ms_per_step": ms_total / steps, "ours_frontend_ms": ms_logmel}
            # parity spot check of the timed configuration against the oracle (same weights, same clips)
            with torch.no_grad():
                xs = x_host[:CPU_SAMPLE_CLIPS].clone()
                y_cpu = O.forward(sd, xs)
                db, mp = model.front_end.logmel_unclamped(xs.to(dev))
                y_gpu = model.encode(db, mp).cpu()
            line["parity_max_abs_err_vs_oracle"] = float((y_gpu - y_cpu).abs().max())
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 300 for 1 s-clip batches, fewer for the large configs)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", choices=list(CONFIGS), default="headline", help="BASELINE.json workload (see module docstring)")
    ap.add_argument("--arch", choices=list(ENCODER_FLOPS), default=None, help="override the config's architecture")
    ap.add_argument("--batch", type=int, default=0, help="override clips per GPU (weak configs) / total clips (c3, c4)")
    ap.add_argument("--precision", choices=["fp32", "bf16"], default=os.environ.get("UITK_PRECISION", "bf16"))
    ap.add_argument("--chunk", type=int, default=512, help="host pipeline chunk (clips)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--diag", choices=["", "nogather", "nogroup"], default="", help="multi-GPU diagnosis: drop the score gather / the max-word exchange (NOT a valid bench line)")
    ap.add_argument("--nccl-gather", action="store_true", help="gather the scores with NCCL instead of PeerGather")
    ap.add_argument("--no-pipeline", action="store_true", help="time the serialised launches only (no BatchPipeline pass)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys of the headline line (10 s front-end, sliding, comparators)")
    args = ap.parse_args()
    if args.impl == "reference":
        if not args.steps:
            args.steps = 5
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
