#!/usr/bin/env python
"""Headline benchmark: UiT 1 s-clip inferences/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arch uit_xs] [--batch 4096] [--precision fp32|bf16]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path (log-mel -> encoder -> scores [-> all-gather when N>1]) over a batch of
`--batch` synthetic 1 s clips PER GPU (weak scaling).  `value` is measured with the batch resident in HBM, `e2e`
through the host-buffer entry point (pinned host waveforms in, host scores out, copies inside the timed region).
See DESIGN.md §Measurement for every field.
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
CPU_SAMPLE_CLIPS = 512


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


def ncu_traffic(kernel: str, arch: str, batch: int):
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/traffic.json), if it was taken
    on this configuration; else None."""
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


def cpu_oracle_throughput(sd, x_cpu, min_seconds: float, max_iters: int):
    """The oracle (torch-CPU restatement of the reference path) on all host threads, bounded sample."""
    import torch
    from oracle import uit_oracle as O
    O.forward(sd, x_cpu[:64])                  # warm-up (thread pool, allocator)
    times, t_start = [], time.perf_counter()
    y = None
    while len(times) < max_iters and (len(times) < 2 or time.perf_counter() - t_start < min_seconds):
        t0 = time.perf_counter()
        y = O.forward(sd, x_cpu)
        times.append(time.perf_counter() - t0)
    return x_cpu.shape[0] / statistics.median(times), torch.get_num_threads(), y, times


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is pure Python
    and /root/reference does not exist on the GPU box, so this times the oracle port (oracle/uit_oracle.py: the
    same ATen CPU ops in the same order, pinned bit-exact/2e-6 against the real reference by
    tests/golden/generate_golden.py) with all host threads, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import uit_mobile_b200 as U
    from oracle import uit_oracle as O
    torch.manual_seed(0)
    model = getattr(U.models, args.arch)(outputdim=537, target_length=102)      # parameter holders on CPU
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = (0.1 * torch.randn(CPU_SAMPLE_CLIPS, 16000, generator=g)).clamp_(-1, 1)
    cores = torch.get_num_threads()
    for _ in range(args.warmup):
        O.forward(sd, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.forward(sd, x)
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE_CLIPS * args.steps / dt
    sample = f"{CPU_SAMPLE_CLIPS} synthetic 1 s clips per step (bounded sample of the {args.batch}-clip workload), fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": f"{args.arch} 1s-clip inferences/sec", "value": value, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.arch} batched inference, synthetic 1 s clips (16 kHz), random-init weights",
                   "arch": args.arch, "batch_per_gpu": args.batch, "clip_samples": 16000},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import uit_mobile_b200 as U
    from uit_mobile_b200 import _native as N
    from uit_mobile_b200 import sharding
    from uit_mobile_b200.pipeline import HostPipeline

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

    torch.manual_seed(0)                                        # identical random-init weights on every rank
    model = getattr(U.models, args.arch)(outputdim=537, target_length=102, precision=args.precision).to(dev).eval()
    if world > 1:
        model.process_group = dist.group.WORLD
    B = args.batch
    total = B * world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = (0.1 * torch.randn(B, 16000, generator=g, device=dev)).clamp_(-1, 1)
    x_host = x.cpu().pin_memory()

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

    # ---- device-resident step, with per-phase events on the launching stream
    pending = [None]      # in-flight all-gather of the previous step (NCCL stream): overlaps the next step's kernels

    def step(marks=None):
        with torch.no_grad():
            db, mp = model.front_end.logmel_unclamped(x)
            if marks is not None:
                marks[1].record()
            if world > 1:
                sharding.allreduce_max_word(mp)
            probs = model.encode(db, mp)
            if marks is not None:
                marks[2].record()
            if world > 1:
                if pending[0] is not None:
                    pending[0][1].wait()
                out = torch.empty((total, probs.shape[1]), dtype=probs.dtype, device=dev)
                pending[0] = (out, dist.all_gather_into_tensor(out, probs, async_op=True), probs)
                probs = out
        return probs

    def drain():
        if pending[0] is not None:
            pending[0][1].wait()
            pending[0] = None

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()
    quiet.__exit__()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.uitk_kernel_launches()
    marks = [[ev(), ev(), ev()] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    for i in range(args.steps):
        marks[i][0].record()
        probs = step(marks[i])
    drain()                 # the last all-gather completes inside the timed region
    e1.record()
    barrier()
    launches = lib.uitk_kernel_launches() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_logmel = statistics.mean(m[0].elapsed_time(m[1]) for m in marks)
    ms_encoder = statistics.mean(m[1].elapsed_time(m[2]) for m in marks)
    value = total * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host in, host scores out)
    pipe = HostPipeline(model, B, 16000, chunk=args.chunk)
    for _ in range(3):
        pipe(x_host)
    barrier()
    e2, e3 = ev(), ev()
    e2.record()
    for _ in range(args.steps):
        out_host = pipe(x_host)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    e2e_value = total * args.steps / (ms_e2e * 1e-3)
    e2e_ok = bool(torch.equal(out_host.to(dev), probs[rank * B:(rank + 1) * B] if world > 1 else probs))

    # ---- the same through 16-bit PCM host buffers (separate, labelled mode: SURVEY §8f n3; halves the H2D bytes)
    pcm_host = (x_host * 32767.0).round().to(torch.int16).pin_memory()
    pipe16 = HostPipeline(model, B, 16000, chunk=args.chunk, dtype=torch.int16)
    for _ in range(3):
        pipe16(pcm_host)
    barrier()
    e4, e5 = ev(), ev()
    e4.record()
    for _ in range(args.steps):
        pipe16(pcm_host)
    e5.record()
    barrier()
    ms_e2e16 = max_over_ranks(e4.elapsed_time(e5))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops = ENCODER_FLOPS[args.arch]
    enc_tflops = B * flops / (ms_encoder * 1e-3) / 1e12
    fe_gbs = B * LOGMEL_BYTES_1S / (ms_logmel * 1e-3) / 1e9
    line = {
        "metric": f"{args.arch} 1s-clip inferences/sec", "value": value, "unit": "clips/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.arch} batched inference, {B} synthetic 1 s clips (16 kHz) per GPU, random-init weights",
                   "arch": args.arch, "batch_per_gpu": B, "global_batch": total, "clip_samples": 16000,
                   "precision": args.precision, "parallelism": f"clips sharded x{world}" if world > 1 else "single GPU",
                   "l2": f"inputs ({B * 64000 / 1e6:.0f} MB/step) larger than the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                "ms_per_step": ms_e2e / args.steps, "chunk": pipe.chunk, "matches_device_path": e2e_ok},
        "e2e_int16_pcm": {"value": total * args.steps / (ms_e2e16 * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": pipe16.h2d_bytes,
                          "d2h_bytes_per_step": pipe16.d2h_bytes, "ms_per_step": ms_e2e16 / args.steps,
                          "note": "same path fed 16-bit PCM host buffers (x = pcm/32768 in-kernel); not the headline"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "encoder (uitk_encoder)", "bound": "tensor", "achieved": enc_tflops,
                     "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": enc_tflops / peaks["bf16_tflops_sustained"],
                     "traffic": ncu_traffic("encoder_tc_kernel", args.arch, B) if args.precision == "bf16" else None,
                     "traffic_unit": "bytes/launch (dram read+write, ncu)", "peak_source": peaks["source"] + " (bf16 cuBLAS, sustained)",
                     "ms_per_launch": ms_encoder, "flops_per_clip": flops},
        "roofline_frontend": {"kernel": "logmel_kernel", "bound": "hbm", "achieved": fe_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": fe_gbs / peaks["hbm_gbs"], "traffic": ncu_traffic("logmel_kernel", args.arch, B),
                              "traffic_unit": "bytes/launch (dram read+write, ncu)", "peak_source": peaks["source"],
                              "ms_per_launch": ms_logmel, "bytes_per_clip": LOGMEL_BYTES_1S},
    }
    # BASELINE config 4 (front-end alone on 10 s clips), bounded to 1024 clips so that the default run stays short
    with torch.no_grad():
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
        bytes10 = 4 * 160000 + 4 * 64 * 1001
        gbs10 = n10 * bytes10 / (ms10 * 1e-3) / 1e9
        line["frontend_10s"] = {"workload": f"log-mel front-end alone, {n10} synthetic 10 s clips (BASELINE config 4, bounded)",
                                "clips_per_s": n10 / (ms10 * 1e-3), "achieved": gbs10, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                "frac": gbs10 / peaks["hbm_gbs"], "bytes_per_clip": bytes10, "ms_per_launch": ms10}
        del x10
        # BASELINE config 5 shape, one GPU's share (57.6 M samples / 8 GPUs): sliding 1 s windows, hop 1600, over one stream
        n_s, hop_s = 7_200_000, 1600
        stream = (0.1 * torch.randn(n_s, generator=g, device=dev)).clamp_(-1, 1)
        n_win = (n_s - 16000) // hop_s + 1

        def per_window():       # every window runs the whole front-end (windows read in place, row stride = hop)
            db_w, mp_w = model.front_end.logmel_unclamped(stream, ld=hop_s, B=n_win, L=16000)
            return model.encode(db_w, mp_w)

        ms_sl = {}
        def shared_stft():      # interior STFT frames computed once for the stream (uitk_logmel_sliding); rank-local: no collective,
            db_w, mp_w = model.front_end.logmel_sliding(stream, 16000, hop_s)      # the other ranks have already left
            return model.encode(db_w, mp_w)

        if world == 1:          # the public entry point (it all-reduces the max word when the model is sharded)
            shared_stft = lambda: model.forward_sliding(stream, hop=hop_s)
        for name, fn in (("per_window_frontend", per_window), ("shared_stft", shared_stft)):
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
            "workload": f"{n_win} sliding 1 s windows, hop {hop_s}, over one {n_s}-sample stream (BASELINE config 5, one GPU's share)",
            "windows_per_s": n_win / (ms_sl["shared_stft"][0] * 1e-3), "ms": ms_sl["shared_stft"][0],
            "ms_per_window_frontend": ms_sl["per_window_frontend"][0],
            "bit_identical_to_per_window_path": bool(torch.equal(ms_sl["shared_stft"][1], ms_sl["per_window_frontend"][1]))}
        del stream
    if world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        xs = x_host[:CPU_SAMPLE_CLIPS].clone()
        cps, cores, y_cpu, times = cpu_oracle_throughput(sd, xs, min_seconds=10.0, max_iters=8)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": f"first {CPU_SAMPLE_CLIPS} clips of the step's batch, {len(times)} passes, median"}
        # parity spot check of the timed configuration against the oracle (same weights, same clips)
        with torch.no_grad():
            db, mp = model.front_end.logmel_unclamped(x[:CPU_SAMPLE_CLIPS].contiguous())
            y_gpu = model.encode(db, mp).cpu()
        line["parity_max_abs_err_vs_oracle"] = float((y_gpu - y_cpu).abs().max())
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--arch", choices=list(ENCODER_FLOPS), default="uit_xs")
    ap.add_argument("--batch", type=int, default=4096, help="clips per GPU per step")
    ap.add_argument("--precision", choices=["fp32", "bf16"], default=os.environ.get("UITK_PRECISION", "bf16"))
    ap.add_argument("--chunk", type=int, default=512, help="host pipeline chunk (clips)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
