"""Host-buffer entry point of the hot path: pinned host waveforms in, host scores out.

This is what a caller sitting where the reference's DataLoader/`inference.py` sits uses when the audio is in host
memory (reference: evaluate.py:53-66 moves every batch host->device and reads the scores back).  The batch is cut
into chunks; the H2D copy of chunk i+1 overlaps the log-mel kernel of chunk i on a second stream.  Because the
top-dB cutoff is batch-global (Q2) every chunk maxes into ONE device word and the encoder runs after the last
chunk's front-end; the scores return through one pinned D2H copy.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as N


class HostPipeline:
    def __init__(self, model, max_batch: int, L: int = 16000, chunk: int = 1024, device: Optional[torch.device] = None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise N.UitkError("HostPipeline needs the model on a CUDA device (no CPU fallback)")
        self.max_batch, self.L, self.chunk = max_batch, L, min(chunk, max_batch)
        T = int(N.lib().uitk_num_frames(L))
        dev = self.device
        self.stage = [torch.empty((self.chunk, L), dtype=torch.float32, device=dev) for _ in range(2)]
        self.db = torch.empty((max_batch, 64, T), dtype=torch.float32, device=dev)
        self.max_pow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out_host = torch.empty((max_batch, model.outputdim), dtype=torch.float32).pin_memory()
        self.copy_stream = torch.cuda.Stream(dev)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @torch.no_grad()
    def __call__(self, wav_host: torch.Tensor) -> torch.Tensor:
        """wav_host: pinned float32 [B, L] host tensor.  Returns a pinned host view [B, outputdim]."""
        if wav_host.is_cuda or wav_host.dtype != torch.float32 or wav_host.dim() != 2 or wav_host.shape[1] != self.L:
            raise ValueError(f"expected a host float32 [B, {self.L}] tensor")
        B = wav_host.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds the pipeline capacity {self.max_batch}")
        m = self.model
        main = torch.cuda.current_stream(self.device)
        self.max_pow.zero_()
        self.copy_stream.wait_stream(main)
        free = [None, None]          # compute-done events per staging buffer
        for i, b0 in enumerate(range(0, B, self.chunk)):
            nb = min(self.chunk, B - b0)
            buf = self.stage[i & 1][:nb]
            with torch.cuda.stream(self.copy_stream):
                if free[i & 1] is not None:
                    self.copy_stream.wait_event(free[i & 1])
                buf.copy_(wav_host[b0:b0 + nb], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            main.wait_event(ready)
            m.front_end.logmel_unclamped(buf, out=self.db[b0:b0 + nb], max_pow=self.max_pow)
            done = torch.cuda.Event()
            done.record(main)
            free[i & 1] = done
        if m.process_group is not None:
            torch.distributed.all_reduce(self.max_pow, op=torch.distributed.ReduceOp.MAX, group=m.process_group)
        probs = m.encode(self.db[:B], self.max_pow)
        out = self.out_host[:B]
        out.copy_(probs, non_blocking=True)
        main.synchronize()
        self.h2d_bytes = B * self.L * 4
        self.d2h_bytes = B * m.outputdim * 4
        return out
